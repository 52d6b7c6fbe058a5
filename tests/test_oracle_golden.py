"""Pins the oracle to the reference's own golden vectors / KATs, CPU only."""
import itertools

import pytest

from oracle import coracle as co
from oracle.pyref import (BN254_FR, CURVES, FIELD_MODULUS, PALLAS_FQ, Keccak256Transcript, SplitMix64,
                          from_mont_bytes, keccak256, mont_bytes, to_repr)


def test_keccak_example():
    # src/provider/keccak.rs:279-288
    assert keccak256((0xFFFFFFFF).to_bytes(4, "little")).hex() == \
        "29045a592007d0c246ef02c2223570da9522d0cf0f73282c79a1bc8f0bb2c238"


@pytest.mark.parametrize("p,h1,h2", [
    # src/provider/keccak.rs:241-258 (non-evm): PallasEngine (scalar = Pallas Fq), Bn256EngineKZG (Fr)
    (PALLAS_FQ, "60dba8657186ff1abbeb237854707faf6ea79361546f8aae65a8fbb722c9ca0c",
     "8bb5dcd9f95115fbc178a1e76d04955423610f5788c7ef2ed200611fecfdf60b"),
    (BN254_FR, "0f8d4f359394760435374d3d603ce0e970ea12f7a05e88eccd52d845f4ac542a",
     "6b32523d63dedd6fb51d5dfc127b9d133cad433ea0b38c4627abadd0e4404c10"),
])
def test_keccak_transcript_golden(p, h1, h2):
    """Pins from_uniform (64-byte LE mod p) and to_repr (LE canonical) of the oracle's field code."""
    t = Keccak256Transcript(p, b"test")
    t.absorb_scalar(b"s1", 2)
    t.absorb_scalar(b"s2", 5)
    assert to_repr(t.squeeze(b"c1")).hex() == h1
    t.absorb_scalar(b"s3", 128)
    assert to_repr(t.squeeze(b"c2")).hex() == h2


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_c_field_matches_bigint(fid):
    """C Montgomery arithmetic == Python integers, incl. the identities of
    src/provider/curve_property_tests.rs:40-72."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(100 + fid)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (1 << 255) % p, (1 << 64) - 1]
    pairs = list(itertools.product(edge, edge)) + [(rng.field(p), rng.field(p)) for _ in range(500)]
    A = b"".join(mont_bytes(p, a) for a, _ in pairs)
    B = b"".join(mont_bytes(p, b) for _, b in pairs)
    for op, fn in [(0, lambda a, b: (a + b) % p), (1, lambda a, b: (a - b) % p), (2, lambda a, b: a * b % p),
                   (3, lambda a, b: pow(a, -1, p) if a else 0), (6, lambda a, b: (-a) % p)]:
        out = co.fe_op(fid, op, A, B)
        for i, (a, b) in enumerate(pairs):
            assert from_mont_bytes(p, out[32 * i:32 * i + 32]) == fn(a, b)
    # from_uniform_bytes(x) == int_LE(x) mod p through the generator used by every harness
    s = co.gen_scalars(fid, 77, 64)
    r = SplitMix64(77)
    assert s == b"".join(mont_bytes(p, r.field(p)) for _ in range(64))
    # a * a^-1 == 1 ; (a+b)(a-b) == a^2 - b^2
    inv = co.fe_op(fid, 3, A)
    prod = co.fe_op(fid, 2, A, inv)
    one = mont_bytes(p, 1)
    for i, (a, _) in enumerate(pairs):
        assert prod[32 * i:32 * i + 32] == (one if a else bytes(32))
    lhs = co.fe_op(fid, 2, co.fe_op(fid, 0, A, B), co.fe_op(fid, 1, A, B))
    rhs = co.fe_op(fid, 1, co.fe_op(fid, 2, A, A), co.fe_op(fid, 2, B, B))
    assert lhs == rhs


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_group_law(cid):
    """src/provider/curve_property_tests.rs:93-116 on the Python group law + generator order."""
    c = CURVES[cid]
    G = c.gen
    assert c.on_curve(G) and c.mul(c.q, G) is None
    rng = SplitMix64(5 + cid)
    P, Q = c.mul(rng.field(c.q), G), c.mul(rng.field(c.q), G)
    k = rng.field(c.q)
    assert c.add(P, Q) == c.add(Q, P)
    assert c.add(P, c.neg(P)) is None
    assert c.mul(k, c.add(P, Q)) == c.add(c.mul(k, P), c.mul(k, Q))
    assert c.add(P, P) == c.mul(2, P)


def _pack_case(c, scalars, bases):
    return (b"".join(mont_bytes(c.q, s) for s in scalars), b"".join(c.affine_bytes(P) for P in bases))


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_c_msm_equals_definition(cid):
    """msm == naive (src/provider/msm.rs:722-739) and the msm_best edge matrix
    (curve_property_tests.rs:172-218: random, all-equal, alternating 0 / (r-1)) at n that
    crosses the n<=16 naive path and the bucket paths."""
    c = CURVES[cid]
    rng = SplitMix64(31 + cid)
    for n in (0, 1, 8, 16, 17, 100):
        bases = c.bases_arith(max(n, 1))[:n]
        for kind in ("random", "equal", "alt"):
            if kind == "random":
                sc = [rng.field(c.q) for _ in range(n)]
            elif kind == "equal":
                v = rng.field(c.q)
                sc = [v] * n
            else:
                sc = [0 if i % 2 == 0 else c.q - 1 for i in range(n)]
            exp = c.msm_naive(sc, bases)
            S, B = _pack_case(c, sc, bases)
            for fn in (co.msm, co.msm_best, co.msm_naive):
                for nt in (1, 3):
                    assert c.affine_from_bytes(fn(cid, S, B, nt)) == exp, (n, kind, fn.__name__)


def test_c_msm_signed_partition_groups():
    """Scalars that land in each of the 11 groups of msm.rs:237-277 (unit, <=8,16,32,64 bits,
    both signs, large) + identity bases with non-zero scalars (msm.rs:788-811)."""
    c = CURVES[0]
    n = 64
    bases = c.bases_arith(n)
    mags = [1, 200, 60000, (1 << 31) + 5, (1 << 63) + 9, (1 << 100) + 3]
    sc = []
    for i in range(n):
        m = mags[i % len(mags)]
        sc.append(m if (i // len(mags)) % 2 == 0 else c.q - m)
    bases[5] = None
    bases[6] = None
    sc[7] = 0
    exp = c.msm_naive(sc, bases)
    S, B = _pack_case(c, sc, bases)
    assert c.affine_from_bytes(co.msm(0, S, B, 4)) == exp
    # all-identity bases -> identity
    B0 = bytes(64 * n)
    assert co.msm(0, S, B0, 2) == bytes(64)


@pytest.mark.parametrize("bits", [1, 4, 8, 10, 16, 20, 32, 40, 64])
def test_c_msm_small_widths(bits):
    """msm_small == msm for the bit-widths of src/provider/msm.rs:751-774."""
    c = CURVES[2]
    rng = SplitMix64(bits)
    n = 80
    bases = c.bases_arith(n)
    sc = [rng.next() & ((1 << bits) - 1) for _ in range(n)]
    sc[0], sc[1] = 0, 1
    exp = c.msm_naive(sc, bases)
    _, B = _pack_case(c, [], bases)
    assert c.affine_from_bytes(co.msm_small(2, sc, B, -1, 3)) == exp
    assert c.affine_from_bytes(co.msm_small(2, sc, B, bits, 1)) == exp


def test_field_vector_kats():
    """bind_poly_var_top order (top variable = MSB): p=(x1+x2)x3 table [0,0,0,1,0,1,0,2]
    evaluates to 2 at (1,1,1) and binding r=1 three times reaches it
    (src/spartan/polys/multilinear.rs:257-281, 391-405)."""
    fid = 0
    p = FIELD_MODULUS[fid]
    Z = co.field_from_u64(fid, [0, 0, 0, 1, 0, 1, 0, 2])
    one = mont_bytes(p, 1)
    z = Z
    for _ in range(3):
        z = co.bind_top(fid, z, one)
    assert from_mont_bytes(p, z) == 2
    # fold/cross-term against Python integers
    rng = SplitMix64(9)
    n = 33
    v = [[rng.field(p) for _ in range(n)] for _ in range(5)]
    u = rng.field(p)
    pk = [b"".join(mont_bytes(p, x) for x in col) for col in v]
    t = co.cross_term(fid, pk[0], pk[1], pk[2], pk[3], None, mont_bytes(p, u))
    t2 = co.cross_term(fid, pk[0], pk[1], pk[2], pk[3], pk[4], mont_bytes(p, u))
    for i in range(n):
        e = (v[0][i] * v[1][i] - u * v[2][i] - v[3][i]) % p
        assert from_mont_bytes(p, t[32 * i:32 * i + 32]) == e
        assert from_mont_bytes(p, t2[32 * i:32 * i + 32]) == (e - v[4][i]) % p
    ax = co.axpy(fid, pk[0], pk[1], mont_bytes(p, u))
    for i in range(n):
        assert from_mont_bytes(p, ax[32 * i:32 * i + 32]) == (v[0][i] + u * v[1][i]) % p


def test_sumcheck_kats_and_c_vs_python():
    """Reference KATs through the C oracle (eq.rs:88-104, multilinear.rs:257-281,327-347,
    sparse.rs:452-465, hyperkzg.rs:1265-1327) and the C sum-check forms against the Python
    EqSumCheckInstance restatement."""
    from oracle import pyref
    fid = 0
    p = FIELD_MODULUS[fid]
    ints = lambda b: [from_mont_bytes(p, b[i:i + 32]) for i in range(0, len(b), 32)]
    assert ints(co.eq_table(fid, co.field_from_u64(fid, [1, 0, 1]))) == [0, 0, 0, 0, 0, 1, 0, 0]
    assert ints(co.mle_eval(fid, co.field_from_u64(fid, [0, 0, 0, 1, 0, 1, 0, 2]), co.field_from_u64(fid, [1, 1, 1]))) == [2]
    assert ints(co.mle_eval(fid, co.field_from_u64(fid, [8, 8, 8, 8]), co.field_from_u64(fid, [3, 4]))) == [8]
    assert ints(co.mle_eval(fid, co.field_from_u64(fid, [1, 2, 1, 4]), co.field_from_u64(fid, [4, 3]))) == [28]
    out = co.spmv(fid, co.field_from_u64(fid, [2, 7, 3, 4]), [1, 2, 2, 0], [0, 2, 3, 4], co.field_from_u64(fid, [1, 2, 3]))
    assert ints(out) == [25, 9, 4]
    rng = SplitMix64(12)
    l = 6
    n = 1 << l
    A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
    taus = [rng.field(p) for _ in range(l)]
    pk = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    # eq table (C, doubling) == python; MLE evaluate
    assert ints(co.eq_table(fid, pk(taus))) == pyref.eq_evals(p, taus)
    assert ints(co.mle_eval(fid, pk(A), pk(taus))) == [pyref.mle_evaluate(p, A, taus)]
    # t(0), t(inf) of round 1 and a later round of the eq-instance, via the C form 4 with the tables
    # the Python instance selects
    eq = pyref.EqSumCheckInstance(p, taus)
    for _ in range(l):
        L, R, sh = eq.tables()
        h = len(A) // 2
        t0 = sum((A[i] * B[i] - C[i]) * eq.factor(i) for i in range(h)) % p
        tinf = sum((A[h + i] - A[i]) * (B[h + i] - B[i]) * eq.factor(i) for i in range(h)) % p
        got = co.sc_eval(fid, 4, pk(A), pk(B), pk(C), pk(L) if L else None, pk(R), sh)
        assert ints(got) == [t0, tinf]
        r = rng.field(p)
        assert ints(co.bind_top(fid, pk(A), mont_bytes(p, r))) == pyref.bind_top(p, A, r)
        A, B, C = pyref.bind_top(p, A, r), pyref.bind_top(p, B, r), pyref.bind_top(p, C, r)
        eq.bound(r)


def test_adx_and_portable_multipliers_agree():
    """oracle.c has two Montgomery multipliers (mulx/adcx/adox inline assembly, selected when the CPU has BMI2 + ADX, and
    the portable unsigned __int128 form): the same field products and the same MSM, bit for bit, in a second
    interpreter with ORACLE_NO_ADX=1.  (The tests above run on whichever this host selects.)"""
    import hashlib
    import os
    import subprocess
    import sys
    prog = ("import sys, hashlib; sys.path.insert(0, %r)\n"
            "from oracle import coracle as co\n"
            "h = hashlib.sha256()\n"
            "for fid in range(4):\n"
            "    A, B = co.gen_scalars(fid, 5, 4096), co.gen_scalars(fid, 6, 4096)\n"
            "    h.update(co.fe_op(fid, 2, A, B)); h.update(co.fe_op(fid, 3, A))\n"
            "for cid in range(4):\n"
            "    h.update(co.msm(cid, co.gen_scalars({0: 0, 1: 1, 2: 3, 3: 2}[cid], 9, 3000), co.gen_bases(cid, 3000)))\n"
            "print(h.hexdigest())\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("0", "1"):
        env = dict(os.environ, ORACLE_NO_ADX=flag)
        outs.append(subprocess.check_output([sys.executable, "-c", prog], env=env, text=True).strip())
    assert outs[0] == outs[1] and len(outs[0]) == 64
