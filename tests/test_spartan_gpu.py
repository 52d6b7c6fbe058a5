"""Parity of the sum-check / MLE / HyperKZG / SpMV kernels (through the C ABI) against the oracle,
with the reference's own KATs (sparse.rs:452-465, multilinear.rs:257-281, eq.rs:88-104,
hyperkzg.rs:1265-1327) and full prover-message equality for the sum-check round loops."""
import pytest

from oracle.pyref import (CURVES, FIELD_MODULUS, Keccak256Transcript, SplitMix64, from_mont_bytes, mont_bytes,
                          prove_cubic_with_three_inputs, prove_quad_prod)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp(b200):
    from nova_b200 import spartan
    return spartan


def ints(p, b):
    return [from_mont_bytes(p, b[i:i + 32]) for i in range(0, len(b), 32)]


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


# ---------------------------------------------------------------- SpMV ------------------------
def test_spmv_kat(sp, oracle):
    """[[0,2,7],[0,0,3],[4,0,0]] * [1,2,3] = [25,9,4]  (sparse.rs:452-465)"""
    fid = 0
    p = FIELD_MODULUS[fid]
    m = sp.SparseMatrix(fid, oracle.field_from_u64(fid, [2, 7, 3, 4]), [1, 2, 2, 0], [0, 2, 3, 4], 3)
    assert ints(p, m.multiply_vec(oracle.field_from_u64(fid, [1, 2, 3]))) == [25, 9, 4]


def _random_csr(rng, p, rows, cols, max_row):
    """mixed coefficient classes incl. empty rows (sparse.rs:488-544)."""
    data, indices, indptr = [], [], [0]
    pool = [1, p - 1, 2, 3, 4, 5, 6, 7, p - 2, p - 3, p - 7, 8, p - 8, 0]
    for r in range(rows):
        k = 0 if r % 17 == 3 else rng.next() % (max_row + 1)
        for _ in range(k):
            sel = rng.next() % 20
            data.append(pool[sel] if sel < len(pool) else rng.field(p))
            indices.append(rng.next() % cols)
        indptr.append(len(indices))
    return data, indices, indptr


@pytest.mark.parametrize("fid", [0, 3])
@pytest.mark.parametrize("rows,cols", [(1, 1), (100, 100), (5000, 3001), (70000, 70000)])
def test_spmv_random(sp, oracle, fid, rows, cols):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(rows + fid)
    data, indices, indptr = _random_csr(rng, p, rows, cols, 6)
    d = pack(p, data)
    z = oracle.gen_scalars(fid, 5, cols)
    m = sp.SparseMatrix(fid, d, indices, indptr, cols)
    assert m.multiply_vec(z) == oracle.spmv(fid, d, indices, indptr, z)


def test_eval_table_sparse_and_gather(sp, oracle):
    """compute_eval_table_sparse (spartan/mod.rs:497-534) as a transposed product, and the
    L_row / L_col gathers of ppsnark.rs:236-250."""
    fid = 0
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(77)
    rows, cols = 4000, 3000
    data, idx, ptr = _random_csr(rng, p, rows, cols, 5)
    d = pack(p, data)
    m = sp.SparseMatrix(fid, d, idx, ptr, cols)
    rx = oracle.gen_scalars(fid, 3, rows)
    for out_len in (cols, 2 * cols, 4096):
        assert m.multiply_transpose(rx, out_len) == oracle.spmv_t(fid, d, idx, ptr, rx, out_len)
    table = oracle.gen_scalars(fid, 4, 1 << 12)
    ix = [rng.next() % (1 << 12) for _ in range(5000)]
    got = sp.gather(table, ix)
    assert got == b"".join(table[32 * i:32 * i + 32] for i in ix)


def test_r1cs_multiply_vec_and_pair(sp, oracle):
    """multiply_vec_pair == 2 x multiply_vec (r1cs/mod.rs:1500-1526) and the tiny cubic R1CS
    x^3 + x + 5 = y (r1cs/mod.rs:1349-1413): rows satisfied <=> Az o Bz == Cz."""
    fid = 0
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(8)
    rows, cols = 3000, 2500
    mats, raw = [], []
    for _ in range(3):
        data, idx, ptr = _random_csr(rng, p, rows, cols, 4)
        raw.append((pack(p, data), idx, ptr))
        mats.append(sp.SparseMatrix(fid, raw[-1][0], idx, ptr, cols))
    S = sp.R1CSShape(*mats)
    z1, z2 = oracle.gen_scalars(fid, 1, cols), oracle.gen_scalars(fid, 2, cols)
    single = S.multiply_vec(z1)
    for k in range(3):
        assert single[k] == oracle.spmv(fid, raw[k][0], raw[k][1], raw[k][2], z1)
    p1, p2 = S.multiply_vec_pair(z1, z2)
    assert p1 == single and p2 == S.multiply_vec(z2)
    with pytest.raises(ValueError):  # InvalidWitnessLength, r1cs/mod.rs:411-413
        S.multiply_vec(z1[:-32])
    # tiny cubic: variables z = (x, x^2, x^3, x^3+x, 1, y)   constraints as in r1cs/mod.rs:1349-1413
    x = 3
    zv = [x, x * x, x ** 3, x ** 3 + x, 1, x ** 3 + x + 5]
    one = 4
    A = ([1, 1, 1, 1, 1, 5], [0, 1, 2, 0, 3, one], [0, 1, 2, 4, 6])
    B = ([1, 1, 1, 1], [0, 0, one, one], [0, 1, 2, 3, 4])
    C = ([1, 1, 1, 1], [1, 2, 3, 5], [0, 1, 2, 3, 4])
    ms = [sp.SparseMatrix(fid, oracle.field_from_u64(fid, d), i, ptr, 6) for d, i, ptr in (A, B, C)]
    az, bz, cz = (ints(p, v) for v in sp.R1CSShape(*ms).multiply_vec(oracle.field_from_u64(fid, zv)))
    assert [a * b % p for a, b in zip(az, bz)] == cz


# ---------------------------------------------------------------- eq / MLE --------------------
def test_eq_and_mle_kats(sp, oracle):
    fid = 0
    p = FIELD_MODULUS[fid]
    t = sp.eq_evals_from_points(fid, oracle.field_from_u64(fid, [1, 0, 1]))  # eq.rs:88-104
    assert ints(p, t) == [0, 0, 0, 0, 0, 1, 0, 0]
    Z = oracle.field_from_u64(fid, [0, 0, 0, 1, 0, 1, 0, 2])  # multilinear.rs:257-281
    assert ints(p, sp.evaluate_with(fid, Z, oracle.field_from_u64(fid, [1, 1, 1]))) == [2]
    Z = oracle.field_from_u64(fid, [8, 8, 8, 8])  # multilinear.rs:327-347
    assert ints(p, sp.evaluate_with(fid, Z, oracle.field_from_u64(fid, [3, 4]))) == [8]
    Z = oracle.field_from_u64(fid, [1, 2, 1, 4])  # hyperkzg.rs:1317-1327: P(4,3) = 28
    assert ints(p, sp.evaluate_with(fid, Z, oracle.field_from_u64(fid, [4, 3]))) == [28]
    Z = oracle.field_from_u64(fid, [1, 2, 2, 4])  # hyperkzg.rs:1265-1313
    for pt, ev in (((0, 0), 1), ((0, 1), 2), ((1, 1), 4), ((0, 2), 3), ((2, 2), 9)):
        assert ints(p, sp.evaluate_with(fid, Z, oracle.field_from_u64(fid, list(pt)))) == [ev]


@pytest.mark.parametrize("fid", [0, 2])
@pytest.mark.parametrize("ell", [0, 1, 2, 7, 10, 11, 15, 18])
def test_eq_table_and_evaluate(sp, oracle, fid, ell):
    r = oracle.gen_scalars(fid, 40 + ell, ell)
    assert sp.eq_evals_from_points(fid, r) == oracle.eq_table(fid, r)
    Z = oracle.gen_scalars(fid, 50 + ell, 1 << ell)
    assert sp.evaluate_with(fid, Z, r) == oracle.mle_eval(fid, Z, r)


# ---------------------------------------------------------------- sum-check forms --------------
@pytest.mark.parametrize("fid", [0, 3])
@pytest.mark.parametrize("form", list(range(11)))
@pytest.mark.parametrize("log_len", [1, 6, 13])
def test_sc_forms(b200, sp, oracle, fid, form, log_len):
    import ctypes
    from nova_b200.native import check, lib
    n = 1 << log_len
    A, B, C = (oracle.gen_scalars(fid, 7 * form + k + log_len, n) for k in range(3))
    count = n if form == 10 else n // 2
    for split in (True, False):
        if form < 4 and not split:
            continue
        bits = max(count.bit_length() - 1, 0)
        shift = bits // 2 if split else 0
        if form >= 4 and split:
            eql = oracle.gen_scalars(fid, 91, max(count >> shift, 1))
            eqr = oracle.gen_scalars(fid, 92, 1 << shift)
        elif form >= 4:
            eql, eqr = None, oracle.gen_scalars(fid, 93, count)
        else:
            eql = eqr = None
        out = ctypes.create_string_buffer(96)
        buf = lambda b: ctypes.create_string_buffer(b, len(b)) if b is not None else None
        check(lib().b200_sc_eval(fid, form, buf(A), buf(B), buf(C), n, buf(eql), len(eql) // 32 if eql else 0,
                                 buf(eqr), len(eqr) // 32 if eqr else 0, shift, out))
        exp = oracle.sc_eval(fid, form, A, B, C, eql, eqr, shift)
        assert out.raw[:len(exp)] == exp, (form, split)


@pytest.mark.parametrize("fid", [0, 3])
def test_sumcheck_provers_match_reference_restatement(sp, oracle, fid):
    """The whole round loop — device sums/binds + host claim derivation + Keccak transcript — emits
    the same compressed polynomials, challenges and final evaluations as the big-integer
    restatement of sumcheck.rs:199-242 and :446-507 (incl. the tau = 0 fall-back)."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(21 + fid)
    for l, zero_tau in ((1, False), (2, False), (7, False), (10, False), (6, True)):
        n = 1 << l
        A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
        taus = [rng.field(p) for _ in range(l)]
        if zero_tau:
            taus[0] = 0
            taus[3] = 0
        claim = rng.field(p)  # the prover does not need a true claim to be deterministic
        exp = prove_cubic_with_three_inputs(p, claim, taus, A, B, C, Keccak256Transcript(p, b"sc"))
        got = sp.SumcheckProof.prove_cubic_with_three_inputs(fid, claim, taus, pack(p, A), pack(p, B), pack(p, C),
                                                             Keccak256Transcript(p, b"sc"))
        assert got == exp, (l, zero_tau)
        exp = prove_quad_prod(p, claim, l, A, B, Keccak256Transcript(p, b"q"))
        got = sp.SumcheckProof.prove_quad_prod(fid, claim, l, pack(p, A), pack(p, B), Keccak256Transcript(p, b"q"))
        assert got == exp


def test_prove_batch_eval_matches_restatement(sp, oracle):
    """SumcheckProof::prove_batch_eval (sumcheck.rs:251-351): instances of different sizes (the
    smaller ones start late with replicated constants), eq instance of the one-input form incl. the
    tau = 0 fall-back; prover messages equal the big-integer restatement."""
    from oracle.pyref import prove_batch_eval
    fid = 0
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(99)
    num_rounds = [9, 6, 9, 1]
    polys = [[rng.field(p) for _ in range(1 << nr)] for nr in num_rounds]
    eq_points = [[rng.field(p) for _ in range(nr)] for nr in num_rounds]
    eq_points[1][2] = 0  # tau = 0 in the middle of the smaller instance
    claims = [rng.field(p) for _ in num_rounds]
    coeffs = [rng.field(p) for _ in num_rounds]
    exp = prove_batch_eval(p, claims, num_rounds, polys, eq_points, coeffs, Keccak256Transcript(p, b"be"))
    got = sp.SumcheckProof.prove_batch_eval(fid, claims, num_rounds, [pack(p, P) for P in polys], eq_points, coeffs,
                                            Keccak256Transcript(p, b"be"))
    assert got == exp


# ---------------------------------------------------------------- batch invert / rlc ----------
@pytest.mark.parametrize("n", [1, 31, 32, 33, 4096, (1 << 15) + 5])
def test_batch_invert(sp, oracle, n):
    """parallel == serial at 2^15+5 (spartan/mod.rs:536-560); zero -> InternalError."""
    fid = 0
    v = oracle.gen_scalars(fid, n, n)
    assert sp.batch_invert(fid, v) == oracle.batch_invert(fid, v)
    bad = v[:32 * (n // 2)] + bytes(32) + v[32 * (n // 2) + 32:]
    with pytest.raises(ValueError):
        sp.batch_invert(fid, bad)
    assert oracle.batch_invert(fid, bad) is None


def test_rlc_diff_sizes(sp, oracle):
    fid = 0
    lens = [1 << 10, 1 << 9, 1 << 10, 37, 0, 1]
    polys = [oracle.gen_scalars(fid, 60 + i, m) for i, m in enumerate(lens)]
    coeffs = oracle.gen_scalars(fid, 70, len(lens))
    assert sp.rlc(fid, polys, coeffs, 1 << 10) == oracle.rlc(fid, polys, coeffs, 1 << 10)


# ---------------------------------------------------------------- HyperKZG pieces --------------
@pytest.mark.parametrize("n", [2, 3, 64, 65, 128, 4097, 1 << 16, (1 << 20) + 3])
def test_poly_eval_div_fold(sp, oracle, n):
    fid = 0
    f = oracle.gen_scalars(fid, n, n)
    us = oracle.gen_scalars(fid, 3, 3)
    assert sp.poly_eval(fid, f, us) == oracle.poly_eval(fid, f, us)
    for k in range(3):
        u = us[32 * k:32 * k + 32]
        assert sp.poly_div(fid, f, u) == oracle.poly_div(fid, f, u)
    if n % 2 == 0:
        assert sp.kzg_fold(fid, f, us[:32]) == oracle.kzg_fold(fid, f, us[:32])


def test_hyperkzg_prove_core(b200, sp, oracle):
    """config-4 shape at test size: every prover message (ell-1 fold commitments, 3 evaluations of
    every fold polynomial, 3 quotient commitments) equals the oracle's (hyperkzg.rs:1076-1116)."""
    cid = 0
    c = CURVES[cid]
    fid = c.scalar_field
    p = c.q
    ell = 10
    n = 1 << ell
    bases = oracle.gen_bases(cid, n)
    ck = b200.CommitmentKey(b200.Curve(cid), bases)
    hat_P = oracle.gen_scalars(fid, 4, n)
    rng = SplitMix64(44)
    x = [rng.field(p) for _ in range(ell)]
    r, q = rng.field(p), rng.field(p)
    com, v, w = sp.hyperkzg_prove_core(cid, ck, hat_P, x, r, q)
    # oracle side
    polys = [hat_P]
    for i in range(ell - 1):
        polys.append(oracle.kzg_fold(fid, polys[i], mont_bytes(p, x[ell - i - 1])))
    aff = lambda b: c.affine_from_bytes(b)
    assert com == [aff(oracle.msm(cid, f, bases[:2 * len(f)])) for f in polys[1:]]
    u = [r, (-r) % p, r * r % p]
    us = pack(p, u)
    assert v == [ints(p, oracle.poly_eval(fid, f, us)) for f in polys]
    Bp = oracle.rlc(fid, polys, pack(p, [pow(q, k, p) for k in range(ell)]), n)
    for t in range(3):
        h = oracle.poly_div(fid, Bp, mont_bytes(p, u[t]))
        assert w[t] == aff(oracle.msm(cid, h, bases[:2 * len(h)]))
    # the last fold polynomial has 2 entries and P_ell = eval is derivable: sanity vs the MLE
    last = polys[-1]
    l0, l1 = ints(p, last)
    assert (l0 + x[0] * (l1 - l0)) % p == from_mont_bytes(p, oracle.mle_eval(fid, hat_P, pack(p, x)))
