"""CPU restatement of the MicroSpartan (ppsnark) prover up to the batched opening claim
(src/spartan/ppsnark.rs:1056-1355).  TEST INFRASTRUCTURE ONLY, like everything under oracle/:
plain Python integers, O(N) loops, for N up to a few hundred.

Parity status: restated from the reference source; the reference itself cannot be executed in this
container (no cargo), so transcripts produced here are "parity unpinned" against a reference
binary.  What the tests pin is (1) internal consistency -- every round polynomial satisfies
s(0) + s(1) = claim and the final claims match direct multilinear evaluations -- and (2) bit-exact
agreement of the CUDA path with this restatement.
"""
from .pyref import (EqSumCheckInstance, Keccak256Transcript, UniPoly, bind_top, commitment_transcript_bytes,
                    eq_evals, mle_evaluate, prove_cubic_with_three_inputs, to_repr, update_claim)


# ---- sumcheck.rs:352-443 ------------------------------------------------------------------
def compute_eval_points_linear(p, A, B):
    h = len(A) // 2
    e0 = sum(A[i] - B[i] for i in range(h)) % p
    einf = sum((2 * A[i] - A[h + i]) - (2 * B[i] - B[h + i]) for i in range(h)) % p
    return e0, einf


def compute_eval_points_quadratic(p, A, B):
    h = len(A) // 2
    e0 = sum(A[i] * B[i] for i in range(h)) % p
    einf = sum((2 * A[i] - A[h + i]) * (2 * B[i] - B[h + i]) for i in range(h)) % p
    return e0, einf


def compute_eval_points_cubic(p, A, B, C):
    h = len(A) // 2
    e0 = cc = einf = 0
    for i in range(h):
        da, db, dc = A[h + i] - A[i], B[h + i] - B[i], C[h + i] - C[i]
        e0 += A[i] * B[i] * C[i]
        cc += da * db * dc
        einf += (A[i] - da) * (B[i] - db) * (C[i] - dc)
    return e0 % p, cc % p, einf % p


def masked_eq_evals(p, r, num_masked_vars):
    """polys/masked_eq.rs:65-76: eq table with the first 2^m entries zeroed."""
    ev = eq_evals(p, r)
    for i in range(1 << num_masked_vars):
        ev[i] = 0
    return ev


def padded(v, n, e=0):
    """ppsnark.rs:41-47."""
    return list(v) + [e] * (n - len(v))


def batch_invert(p, v):
    """spartan/mod.rs:54-145 (value-level: element-wise inverse; zero -> error)."""
    if any(x % p == 0 for x in v):
        raise ZeroDivisionError("batch_invert: zero element")
    return [pow(x, -1, p) for x in v]


# ---- R1CSShapeSparkRepr (ppsnark.rs:113-198, 220-253) -------------------------------------------
class SparkRepr:
    def __init__(self, p, A, B, C, num_cons, num_vars):
        """A, B, C: lists of (row, col, val) in the order the reference's SparseMatrix::iter yields
        (CSR order)."""
        total = len(A) + len(B) + len(C)
        n = max(total, 2 * num_vars, num_cons)
        N = 1
        while N < n:
            N *= 2
        self.p, self.N = p, N
        row, col = [0] * N, [N - 1] * N
        for i, (r, c, _) in enumerate(A + B + C):
            row[i], col[i] = r, c
        self.val_A, self.val_B, self.val_C = [0] * N, [0] * N, [0] * N
        for i, (_, _, v) in enumerate(A):
            self.val_A[i] = v % p
        for i, (_, _, v) in enumerate(B):
            self.val_B[len(A) + i] = v % p
        for i, (_, _, v) in enumerate(C):
            self.val_C[len(A) + len(B) + i] = v % p
        ts_row, ts_col = [0] * N, [0] * N
        for a in row:
            ts_row[a] += 1
        for a in col:
            ts_col[a] += 1
        self.row_idx, self.col_idx = row, col
        self.row, self.col, self.ts_row, self.ts_col = list(row), list(col), ts_row, ts_col
        self.nnz = total

    def evaluation_oracles(self, r_outer, z):
        p, N = self.p, self.N
        mem_row = eq_evals(p, r_outer)
        mem_col = padded(z, N)
        L_row = [mem_row[self.row_idx[i]] for i in range(N)]  # padding rows are 0 -> mem_row[0]
        L_col = [mem_col[self.col_idx[i]] for i in range(N)]  # padding cols are N-1 -> mem_col[N-1]
        return mem_row, mem_col, L_row, L_col


# ---- the three engines (ppsnark.rs:270-786) -----------------------------------------------------
class WitnessBoundSumcheck:
    def __init__(self, p, tau, W_padded, num_vars):
        m = num_vars.bit_length() - 1
        assert m < len(W_padded).bit_length() - 1
        self.p = p
        self.W = list(W_padded)
        self.masked_eq = masked_eq_evals(p, tau, m)

    def initial_claims(self):
        return [0]

    def size(self):
        return len(self.W)

    def evaluation_points(self):
        e0, einf = compute_eval_points_quadratic(self.p, self.masked_eq, self.W)
        return [[e0, 0, einf]]

    def bound(self, r):
        self.W = bind_top(self.p, self.W, r)
        self.masked_eq = bind_top(self.p, self.masked_eq, r)

    def final_claims(self):
        return [[self.W[0], self.masked_eq[0]]]


def memory_compute_oracles(p, r, gamma, mem_row, addr_row, L_row, ts_row, mem_col, addr_col, L_col, ts_col):
    """MemorySumcheckInstance::compute_oracles without the four commitments (ppsnark.rs:372-495):
    returns (poly_vec, aux_poly_vec)."""
    def side(mem, addr, L, ts):
        T = [(mem[i] * gamma + i) % p for i in range(len(mem))]
        W = [(L[i] * gamma + addr[i]) % p for i in range(len(addr))]
        tpr = [(t + r) % p for t in T]
        wpr = [(w + r) % p for w in W]
        inv = batch_invert(p, tpr + wpr)
        t_inv = [inv[i] * ts[i] % p for i in range(len(T))]
        w_inv = inv[len(T):]
        return t_inv, w_inv, tpr, wpr
    tir, wir, tr, wr = side(mem_row, addr_row, L_row, ts_row)
    tic, wic, tc, wc = side(mem_col, addr_col, L_col, ts_col)
    return [tir, wir, tic, wic], [tr, wr, tc, wc]


class MemorySumcheckInstance:
    def __init__(self, p, polys_oracle, polys_aux, rhos, ts_row, ts_col):
        self.p = p
        self.t_inv_row, self.w_inv_row, self.t_inv_col, self.w_inv_col = [list(v) for v in polys_oracle]
        self.t_row, self.w_row, self.t_col, self.w_col = [list(v) for v in polys_aux]
        self.ts_row, self.ts_col = list(ts_row), list(ts_col)
        self.eq = EqSumCheckInstance(p, rhos)
        self.running = [0] * 6
        self.saved = [[0, 0, 0] for _ in range(6)]

    def initial_claims(self):
        return [0] * 6

    def size(self):
        return len(self.w_row)

    def evaluation_points(self):
        p, eq, rc = self.p, self.eq, self.running
        i0r, i3r = compute_eval_points_linear(p, self.t_inv_row, self.w_inv_row)
        i0c, i3c = compute_eval_points_linear(p, self.t_inv_col, self.w_inv_col)
        Tr = eq.evaluation_points_cubic_with_three_inputs(self.t_inv_row, self.t_row, self.ts_row, rc[2])
        Wr = eq.evaluation_points_cubic_with_two_inputs(self.w_inv_row, self.w_row, rc[3])
        Tc = eq.evaluation_points_cubic_with_three_inputs(self.t_inv_col, self.t_col, self.ts_col, rc[4])
        Wc = eq.evaluation_points_cubic_with_two_inputs(self.w_inv_col, self.w_col, rc[5])
        self.saved = [[i0r, 0, i3r], [i0c, 0, i3c], list(Tr), list(Wr), list(Tc), list(Wc)]
        return [list(e) for e in self.saved]

    def bound(self, r):
        p = self.p
        self.running = [update_claim(p, self.running[j], self.saved[j], r) for j in range(6)]
        for name in ("t_row", "t_inv_row", "w_row", "w_inv_row", "ts_row", "t_col", "t_inv_col", "w_col",
                     "w_inv_col", "ts_col"):
            setattr(self, name, bind_top(p, getattr(self, name), r))
        self.eq.bound(r)

    def final_claims(self):
        return [[self.t_inv_row[0], self.w_inv_row[0], self.ts_row[0]],
                [self.t_inv_col[0], self.w_inv_col[0], self.ts_col[0]]]


class InnerBatchedSumcheckInstance:
    def __init__(self, p, claim, L_row, L_col, val, claim_E, r_outer, E):
        self.p, self.claim, self.claim_E = p, claim % p, claim_E % p
        self.L_row, self.L_col, self.val, self.E = list(L_row), list(L_col), list(val), list(E)
        self.eq = EqSumCheckInstance(p, r_outer)
        self.running_E = claim_E % p
        self.saved_E = [0, 0, 0]

    def initial_claims(self):
        return [self.claim, self.claim_E]

    def size(self):
        return len(self.L_row)

    def evaluation_points(self):
        e0, bc, einf = compute_eval_points_cubic(self.p, self.L_row, self.L_col, self.val)
        E0, Eb, Einf = self.eq.evaluation_points_quadratic_with_one_input(self.E, self.running_E)
        self.saved_E = [E0, Eb, Einf]
        return [[e0, bc, einf], [E0, 0, Einf]]

    def bound(self, r):
        p = self.p
        self.running_E = update_claim(p, self.running_E, self.saved_E, r)
        for name in ("L_row", "L_col", "val", "E"):
            setattr(self, name, bind_top(p, getattr(self, name), r))
        self.eq.bound(r)

    def final_claims(self):
        return [[self.L_row[0], self.L_col[0]], [self.E[0]]]


def prove_helper(p, mem, inner, witness, transcript):
    """RelaxedR1CSSNARK::prove_helper (ppsnark.rs:886-983)."""
    assert mem.size() == inner.size() == witness.size()
    claims = mem.initial_claims() + inner.initial_claims() + witness.initial_claims()
    s = transcript.squeeze(b"r")
    coeffs = [pow(s, i, p) for i in range(len(claims))]  # powers(), spartan/mod.rs:40-48
    e = sum(c * k for c, k in zip(claims, coeffs)) % p
    rs, polys = [], []
    for _ in range(mem.size().bit_length() - 1):
        evals = mem.evaluation_points() + inner.evaluation_points() + witness.evaluation_points()
        assert len(evals) == len(claims)
        c0 = sum(evals[i][0] * coeffs[i] for i in range(len(evals))) % p
        cb = sum(evals[i][1] * coeffs[i] for i in range(len(evals))) % p
        ci = sum(evals[i][2] * coeffs[i] for i in range(len(evals))) % p
        poly = UniPoly.from_evals_deg3(p, [c0, (e - c0) % p, cb, ci])
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        mem.bound(r)
        inner.bound(r)
        witness.bound(r)
        e = poly.evaluate(r)
        polys.append(poly.compressed())
    return polys, rs, mem.final_claims(), inner.final_claims(), witness.final_claims()


def scalars_bytes(xs):
    return b"".join(to_repr(x) for x in xs)


def commitments_bytes(Ps):
    return b"".join(commitment_transcript_bytes(P) for P in Ps)


def prove_core(p, commit, S, spark, U, W, vk_digest):
    """ppsnark.rs:1056-1355 up to (and excluding) EE::prove.

    commit(vec) -> affine point or None (identity): the Pedersen/KZG commitment with r = 0.
    S: dict(num_cons, num_vars, A, B, C) with A/B/C lists of (row, col, val) (already padded/regular).
    U: dict(comm_W, comm_E (affine or None), u, X);  W: dict(W, E).
    Returns a dict with every proof field, the batched opening polynomial and its claimed value.
    """
    num_cons, num_vars, N = S["num_cons"], S["num_vars"], spark.N
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    tr.absorb_scalar(b"vk", vk_digest)
    tr.absorb_bytes(b"U", commitments_bytes([U["comm_W"], U["comm_E"]]) + to_repr(U["u"] % p) + scalars_bytes(U["X"]))
    z = list(W["W"]) + [U["u"]] + list(U["X"])

    def spmv(M):
        out = [0] * num_cons
        for (r, c, v) in M:
            out[r] = (out[r] + v * z[c]) % p
        return out
    Az, Bz, Cz = spmv(S["A"]), spmv(S["B"]), spmv(S["C"])
    nro, nri = num_cons.bit_length() - 1, N.bit_length() - 1
    tau = [tr.squeeze(b"t") for _ in range(nro)]
    uCz_E = [(U["u"] * c + e) % p for c, e in zip(Cz, W["E"])]
    sc_outer, r_outer, claims_outer = prove_cubic_with_three_inputs(p, 0, tau, Az, Bz, uCz_E, tr)
    eAz, eBz = claims_outer[0], claims_outer[1]
    eCz = mle_evaluate(p, Cz, r_outer)
    eE_outer = (claims_outer[2] - U["u"] * eCz) % p
    tr.absorb_bytes(b"e", scalars_bytes([eAz, eBz, eCz, eE_outer]))
    r_pad = [tr.squeeze(b"p") for _ in range(nri - nro)]
    r_outer_full = r_pad + r_outer
    factor = 1
    for x in r_pad:
        factor = factor * (1 - x) % p
    E_p, W_p = padded(W["E"], N), padded(W["W"], N)
    mem_row, mem_col, L_row, L_col = spark.evaluation_oracles(r_outer_full, z)
    comm_L_row, comm_L_col = commit(L_row), commit(L_col)
    tr.absorb_bytes(b"e", commitments_bytes([comm_L_row, comm_L_col]))
    c = tr.squeeze(b"c")
    gamma = tr.squeeze(b"g")
    r = tr.squeeze(b"r")
    val = [(a + c * b + c * c * cc) % p for a, b, cc in zip(spark.val_A, spark.val_B, spark.val_C)]
    inner = InnerBatchedSumcheckInstance(p, factor * (eAz + c * eBz + c * c * eCz), L_row, L_col, val,
                                         factor * eE_outer, r_outer_full, E_p)
    mem_oracles, mem_aux = memory_compute_oracles(p, r, gamma, mem_row, spark.row, L_row, spark.ts_row,
                                                  mem_col, spark.col, L_col, spark.ts_col)
    comm_mem = [commit(v) for v in mem_oracles]
    tr.absorb_bytes(b"l", commitments_bytes(comm_mem))
    rho = [tr.squeeze(b"r") for _ in range(nri)]
    mem = MemorySumcheckInstance(p, mem_oracles, mem_aux, rho, spark.ts_row, spark.ts_col)
    wit = WitnessBoundSumcheck(p, r_outer_full, W_p, num_vars)
    sc_inner, r_inner, c_mem, c_inner, c_wit = prove_helper(p, mem, inner, wit, tr)
    ev = {
        "eval_L_row": c_inner[0][0], "eval_L_col": c_inner[0][1], "eval_E": c_inner[1][0],
        "eval_t_plus_r_inv_row": c_mem[0][0], "eval_w_plus_r_inv_row": c_mem[0][1], "eval_ts_row": c_mem[0][2],
        "eval_t_plus_r_inv_col": c_mem[1][0], "eval_w_plus_r_inv_col": c_mem[1][1], "eval_ts_col": c_mem[1][2],
        "eval_W": c_wit[0][0],
    }
    for name, v in (("eval_val_A", spark.val_A), ("eval_val_B", spark.val_B), ("eval_val_C", spark.val_C),
                    ("eval_row", spark.row), ("eval_col", spark.col)):
        ev[name] = mle_evaluate(p, v, r_inner)
    order = ["eval_W", "eval_E", "eval_L_row", "eval_L_col", "eval_val_A", "eval_val_B", "eval_val_C",
             "eval_t_plus_r_inv_row", "eval_row", "eval_w_plus_r_inv_row", "eval_ts_row",
             "eval_t_plus_r_inv_col", "eval_col", "eval_w_plus_r_inv_col", "eval_ts_col"]
    eval_vec = [ev[k] for k in order]
    poly_vec = [W_p, E_p, L_row, L_col, spark.val_A, spark.val_B, spark.val_C, mem_oracles[0], spark.row,
                mem_oracles[1], spark.ts_row, mem_oracles[2], spark.col, mem_oracles[3], spark.ts_col]
    tr.absorb_bytes(b"e", scalars_bytes(eval_vec))
    cb = tr.squeeze(b"c")
    pw = [pow(cb, i, p) for i in range(len(poly_vec))]  # PolyEvalWitness::batch, spartan/mod.rs:232-247
    batched = [sum(pw[k] * poly_vec[k][i] for k in range(len(poly_vec))) % p for i in range(N)]
    batched_eval = sum(pw[k] * eval_vec[k] for k in range(len(eval_vec))) % p
    out = dict(ev)
    out.update(comm_L_row=comm_L_row, comm_L_col=comm_L_col, comm_mem=comm_mem, sc_outer=sc_outer,
               r_outer=r_outer, eval_Az_at_r_outer=eAz, eval_Bz_at_r_outer=eBz, eval_Cz_at_r_outer=eCz,
               eval_E_at_r_outer=eE_outer, sc_inner_batched=sc_inner, r_inner_batched=r_inner,
               batched_poly=batched, batched_eval=batched_eval, batch_challenge=cb, transcript=tr)
    return out


# ---- verifier side (ppsnark.rs:1386-1640), used to pin the restatement above ----------------------
def sumcheck_verify(p, compressed_polys, claim, num_rounds, degree_bound, transcript):
    """SumcheckProof::verify (sumcheck.rs:87-127) with CompressedUniPoly::decompress
    (polys/univariate.rs:161-174)."""
    e, rs = claim % p, []
    assert len(compressed_polys) == num_rounds
    for cp in compressed_polys:
        lin = (e - 2 * cp[0] - sum(cp[1:])) % p
        poly = UniPoly(p, [cp[0], lin] + list(cp[1:]))
        assert len(poly.coeffs) - 1 <= degree_bound
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        e = poly.evaluate(r)
    return e, rs


def eq_evaluate(p, r, rx):
    out = 1
    for a, b in zip(r, rx):
        out = out * (a * b + (1 - a) * (1 - b)) % p
    return out


def masked_eq_evaluate(p, r, m, rx):
    """polys/masked_eq.rs:34-52."""
    split = len(r) - m
    eq_lo = eq_evaluate(p, r[:split], rx[:split])
    eq_hi = eq_evaluate(p, r[split:], rx[split:])
    mask_lo = 1
    for a, b in zip(r[:split], rx[:split]):
        mask_lo = mask_lo * (1 - a) * (1 - b) % p
    return (eq_lo - mask_lo) * eq_hi % p


def identity_evaluate(p, r):
    """polys/identity.rs:21-34."""
    return sum(pow(2, len(r) - 1 - i, p) * r[i] for i in range(len(r))) % p


def sparse_poly_evaluate(p, num_vars, Z, r):
    """polys/multilinear.rs:207-225."""
    assert len(r) == num_vars
    nz = 1
    while nz < len(Z):
        nz *= 2
    nvz = nz.bit_length() - 1
    assert num_vars - 1 - nvz >= 0, "public IO too long for the shape (usize underflow panic in the reference)"
    chis = eq_evals(p, r[num_vars - 1 - nvz:])
    partial = sum(z * c for z, c in zip(Z, chis)) % p
    common = 1
    for i in range(num_vars - 1 - nvz):
        common = common * (1 - r[i]) % p
    return common * partial % p


def verify_core(p, num_cons, num_vars, N, U, vk_digest, proof, holder=None):
    """Re-derives every challenge and checks both sum-check final claims.  Returns True / raises.
    `holder` (optional dict) receives the transcript so that the opening check can continue it."""
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    if holder is not None:
        holder["tr"] = tr
    tr.absorb_scalar(b"vk", vk_digest)
    tr.absorb_bytes(b"U", commitments_bytes([U["comm_W"], U["comm_E"]]) + to_repr(U["u"] % p) + scalars_bytes(U["X"]))
    nro, nri = num_cons.bit_length() - 1, N.bit_length() - 1
    tau = [tr.squeeze(b"t") for _ in range(nro)]
    claim_outer, r_outer = sumcheck_verify(p, proof["sc_outer"], 0, nro, 3, tr)
    eAz, eBz, eCz, eE = (proof[k] for k in ("eval_Az_at_r_outer", "eval_Bz_at_r_outer", "eval_Cz_at_r_outer",
                                             "eval_E_at_r_outer"))
    assert eq_evaluate(p, tau, r_outer) * (eAz * eBz - U["u"] * eCz - eE) % p == claim_outer, "outer sum-check"
    tr.absorb_bytes(b"e", scalars_bytes([eAz, eBz, eCz, eE]))
    r_pad = [tr.squeeze(b"p") for _ in range(nri - nro)]
    r_full = r_pad + r_outer
    factor = 1
    for x in r_pad:
        factor = factor * (1 - x) % p
    tr.absorb_bytes(b"e", commitments_bytes([proof["comm_L_row"], proof["comm_L_col"]]))
    c = tr.squeeze(b"c")
    gamma = tr.squeeze(b"g")
    r = tr.squeeze(b"r")
    tr.absorb_bytes(b"l", commitments_bytes(proof["comm_mem"]))
    rho = [tr.squeeze(b"r") for _ in range(nri)]
    s = tr.squeeze(b"r")
    co = [pow(s, i, p) for i in range(9)]
    claim = (co[6] * factor * (eAz + c * eBz + c * c * eCz) + co[7] * factor * eE) % p
    final, ri = sumcheck_verify(p, proof["sc_inner_batched"], claim, nri, 3, tr)
    g = proof
    rand_eq = eq_evaluate(p, rho, ri)
    eq_ro = eq_evaluate(p, r_full, ri)
    masked = masked_eq_evaluate(p, r_full, num_vars.bit_length() - 1, ri)
    ident = identity_evaluate(p, ri)
    t_row = (ident + gamma * eq_ro + r) % p
    w_row = (g["eval_row"] + gamma * g["eval_L_row"] + r) % p
    l = nri - (2 * num_vars).bit_length() + 1
    fac2 = 1
    for x in ri[:l]:
        fac2 = fac2 * (1 - x) % p
    unpad = ri[l:]
    eval_X = sparse_poly_evaluate(p, len(unpad) - 1, [U["u"]] + list(U["X"]), unpad[1:])
    eval_Z = (g["eval_W"] + fac2 * unpad[0] * eval_X) % p
    t_col = (ident + gamma * eval_Z + r) % p
    w_col = (g["eval_col"] + gamma * g["eval_L_col"] + r) % p
    expected = (co[0] * (g["eval_t_plus_r_inv_row"] - g["eval_w_plus_r_inv_row"])
                + co[1] * (g["eval_t_plus_r_inv_col"] - g["eval_w_plus_r_inv_col"])
                + co[2] * rand_eq * (g["eval_t_plus_r_inv_row"] * t_row - g["eval_ts_row"])
                + co[3] * rand_eq * (g["eval_w_plus_r_inv_row"] * w_row - 1)
                + co[4] * rand_eq * (g["eval_t_plus_r_inv_col"] * t_col - g["eval_ts_col"])
                + co[5] * rand_eq * (g["eval_w_plus_r_inv_col"] * w_col - 1)
                + co[6] * g["eval_L_row"] * g["eval_L_col"] * (g["eval_val_A"] + c * g["eval_val_B"] + c * c * g["eval_val_C"])
                + co[7] * eq_ro * g["eval_E"]
                + co[8] * masked * g["eval_W"]) % p
    assert expected == final, "inner batched sum-check"
    assert ri == proof["r_inner_batched"]
    return True


# ---- test input generator ------------------------------------------------------------------------
def random_instance(p, rng, num_cons, num_vars, num_io, nnz_per_row=2):
    """Random regular shape + satisfying relaxed witness (E absorbs the slack)."""
    ncols = num_vars + 1 + num_io

    def mat():
        out = []
        for r in range(num_cons):
            cols = sorted({rng.next() % ncols for _ in range(nnz_per_row)})
            for c in cols:
                v = [1, p - 1, 2, rng.field(p)][rng.next() % 4]
                out.append((r, c, v))
        return out
    A, B, C = mat(), mat(), mat()
    Wv = [rng.field(p) for _ in range(num_vars)]
    X = [rng.field(p) for _ in range(num_io)]
    u = rng.field(p)
    z = Wv + [u] + X

    def mv(M):
        out = [0] * num_cons
        for (r, c, v) in M:
            out[r] = (out[r] + v * z[c]) % p
        return out
    Az, Bz, Cz = mv(A), mv(B), mv(C)
    E = [(a * b - u * c) % p for a, b, c in zip(Az, Bz, Cz)]
    S = dict(num_cons=num_cons, num_vars=num_vars, A=A, B=B, C=C)
    return S, dict(W=Wv, E=E), u, X


# ---- the whole prover / verifier incl. the evaluation argument (ppsnark.rs:1305-1340, 1603-1655) -----------
EVAL_ORDER = ["eval_W", "eval_E", "eval_L_row", "eval_L_col", "eval_val_A", "eval_val_B", "eval_val_C",
              "eval_t_plus_r_inv_row", "eval_row", "eval_w_plus_r_inv_row", "eval_ts_row",
              "eval_t_plus_r_inv_col", "eval_col", "eval_w_plus_r_inv_col", "eval_ts_col"]


def shape_commitments(commit, spark):
    """R1CSShapeSparkCommitment (ppsnark.rs:200-215): commitments to the seven preprocessed vectors."""
    return {k: commit(getattr(spark, k)) for k in ("val_A", "val_B", "val_C", "row", "col", "ts_row", "ts_col")}


def comm_vec_of(U, S_comm, proof):
    cm = proof["comm_mem"]
    return [U["comm_W"], U["comm_E"], proof["comm_L_row"], proof["comm_L_col"], S_comm["val_A"], S_comm["val_B"],
            S_comm["val_C"], cm[0], S_comm["row"], cm[1], S_comm["ts_row"], cm[2], S_comm["col"], cm[3], S_comm["ts_col"]]


def batch_commitment(p, curve, comm_vec, c):
    """PolyEvalInstance::batch, the commitment part (spartan/mod.rs:346-368): sum_i c^i C_i."""
    C = None
    for i, cm in enumerate(comm_vec):
        C = curve.add(C, curve.mul(pow(c, i, p), cm))
    return C


def prove(p, curve, cid, srs, commit, S, spark, U, W, vk_digest):
    """RelaxedR1CSSNARK::prove of ppsnark.rs incl. EE::prove (HyperKZG, transcript-driven)."""
    from . import hyperkzg_ref as hk
    from .pyref import mont_bytes
    out = prove_core(p, commit, S, spark, U, W, vk_digest)
    hat_P = b"".join(mont_bytes(p, v) for v in out["batched_poly"])
    out["eval_arg"] = hk.prove(cid, srs, hat_P, out["r_inner_batched"], out["transcript"])
    return out


def verify(p, curve, cid, tau, num_cons, num_vars, N, U, S_comm, vk_digest, proof) -> bool:
    from . import hyperkzg_ref as hk
    holder = {}
    verify_core(p, num_cons, num_vars, N, U, vk_digest, proof, holder)
    tr = holder["tr"]
    eval_vec = [proof[k] for k in EVAL_ORDER]
    tr.absorb_bytes(b"e", scalars_bytes(eval_vec))
    c = tr.squeeze(b"c")
    C = batch_commitment(p, curve, comm_vec_of(U, S_comm, proof), c)
    e = sum(pow(c, i, p) * v for i, v in enumerate(eval_vec)) % p
    return hk.verify(cid, tau, C, proof["r_inner_batched"], e, proof["eval_arg"], tr)
