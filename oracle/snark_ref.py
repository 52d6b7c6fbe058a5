"""CPU restatement of spartan::snark::RelaxedR1CSSNARK::prove up to the evaluation argument
(src/spartan/snark.rs:113-256) and of the matching part of ::verify
(snark.rs:259-396).  TEST INFRASTRUCTURE ONLY: plain Python integers, O(N) loops, small N.

prove_core  = prove minus EE::prove: outer sum-check (cubic, eq(tau)(Az Bz - (u Cz + E))), claims,
              inner sum-check (quadratic, (A + rB + r^2 C)(r_x, .) * z), eval_W, batch_eval_reduce
              (spartan/mod.rs:377-432) down to the single claim (C, x, e) and the batched polynomial
              handed to EE::prove.
verify_core = verify minus EE::verify: re-derives every challenge from the transcript and checks both
              final sum-check claims against the R1CS matrices and the batch-evaluation claim
              (spartan/mod.rs:436-480).

Parity status: the reference cannot be executed here (no cargo); the prover restatement is pinned by
the verifier restatement -- equations it is not built from -- on satisfying instances, and must be
rejected when the witness, a claim or a round polynomial is altered (tests/test_snark_oracle.py).
The CUDA path is then compared with prove_core message by message.
"""
from .ppsnark_ref import (commitments_bytes, eq_evaluate, scalars_bytes, sparse_poly_evaluate, sumcheck_verify)
from .pyref import (Keccak256Transcript, eq_evals, mle_evaluate, prove_batch_eval, prove_cubic_with_three_inputs,
                    prove_quad_prod, to_repr)


def _log2(n):
    assert n > 0 and n & (n - 1) == 0, "regular shape: powers of two (r1cs/mod.rs is_regular_shape)"
    return n.bit_length() - 1


def _start_transcript(p, U, vk_digest):
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")  # snark.rs:125
    tr.absorb_scalar(b"vk", vk_digest)                # :128
    tr.absorb_bytes(b"U", commitments_bytes([U["comm_W"], U["comm_E"]]) + to_repr(U["u"] % p)
                    + scalars_bytes(U["X"]))          # :129, RelaxedR1CSInstance::to_transcript_bytes
    return tr


def spmv(p, M, z, rows):
    out = [0] * rows
    for (r, c, v) in M:
        out[r] = (out[r] + v * z[c]) % p
    return out


def eval_table_sparse(p, M, rx, width):
    """compute_eval_table_sparse (spartan/mod.rs:497-534): M_evals[col] += rx[row] * val."""
    out = [0] * width
    for (r, c, v) in M:
        out[c] = (out[c] + rx[r] * v) % p
    return out


def batch_eval_reduce(p, curve, u_vec, w_vec, tr):
    """spartan/mod.rs:377-432.  u_vec: [(commitment, x, e)], w_vec: [poly].
    -> (joint (C, x, e), joint polynomial, c, sc_proof, claims_batch_left)."""
    num_rounds = [len(x) for (_, x, _) in u_vec]
    for w, nr in zip(w_vec, num_rounds):
        assert len(w) == 1 << nr
    rho = tr.squeeze(b"r")
    powers = [pow(rho, i, p) for i in range(len(u_vec))]
    claims = [e for (_, _, e) in u_vec]
    xs = [x for (_, x, _) in u_vec]
    sc, r, left = prove_batch_eval(p, claims, num_rounds, w_vec, xs, powers, tr)
    tr.absorb_bytes(b"l", scalars_bytes(left))
    c = tr.squeeze(b"c")
    joint_u = batch_diff_size_instance(p, curve, [cm for (cm, _, _) in u_vec], left, num_rounds, r, c)
    size_max = max(len(w) for w in w_vec)
    joint_w = [sum(pow(c, k, p) * (w[i] if i < len(w) else 0) for k, w in enumerate(w_vec)) % p
               for i in range(size_max)]               # PolyEvalWitness::batch_diff_size, mod.rs:165-222
    return joint_u, joint_w, c, sc, left


def batch_diff_size_instance(p, curve, comms, evals, num_vars, x, s):
    """PolyEvalInstance::batch_diff_size (spartan/mod.rs:304-344)."""
    nmax = len(x)
    e = 0
    C = None
    for i, (cm, ev, nv) in enumerate(zip(comms, evals, num_vars)):
        lag = 1
        for rr in x[:nmax - nv]:
            lag = lag * (1 - rr) % p
        g = pow(s, i, p)
        e = (e + g * lag * ev) % p
        C = curve.add(C, curve.mul(g, cm))
    return C, list(x), e


def prove_core(p, curve, S, U, W, vk_digest):
    """S: dict(num_cons, num_vars, A, B, C) with triplet lists; U: dict(comm_W, comm_E, u, X);
    W: dict(W, E).  Returns every proof field of RelaxedR1CSSNARK except eval_arg, plus the claim and
    polynomial that go to EE::prove."""
    num_cons, num_vars = S["num_cons"], S["num_vars"]
    nrx, nry = _log2(num_cons), _log2(num_vars) + 1
    assert len(U["X"]) < num_vars
    tr = _start_transcript(p, U, vk_digest)
    z = list(W["W"]) + [U["u"] % p] + [x % p for x in U["X"]]
    tau = [tr.squeeze(b"t") for _ in range(nrx)]
    Az, Bz, Cz = (spmv(p, S[k], z, num_cons) for k in "ABC")
    uCz_E = [(U["u"] * c + e) % p for c, e in zip(Cz, W["E"])]
    sc_outer, r_x, claims_outer = prove_cubic_with_three_inputs(p, 0, tau, Az, Bz, uCz_E, tr)
    claim_Az, claim_Bz = claims_outer[0], claims_outer[1]
    claim_Cz = mle_evaluate(p, Cz, r_x)
    eval_E = mle_evaluate(p, W["E"], r_x)
    tr.absorb_bytes(b"claims_outer", scalars_bytes([claim_Az, claim_Bz, claim_Cz, eval_E]))
    r = tr.squeeze(b"r")
    claim_inner_joint = (claim_Az + r * claim_Bz + r * r * claim_Cz) % p
    evals_rx = eq_evals(p, r_x)
    eA, eB, eC = (eval_table_sparse(p, S[k], evals_rx, 2 * num_vars) for k in "ABC")
    poly_ABC = [(a + r * b + r * r * c) % p for a, b, c in zip(eA, eB, eC)]
    poly_z = z + [0] * (2 * num_vars - len(z))
    sc_inner, r_y, _ = prove_quad_prod(p, claim_inner_joint, nry, poly_ABC, poly_z, tr)
    eval_W = mle_evaluate(p, W["W"], r_y[1:])
    tr.absorb_bytes(b"w", scalars_bytes([eval_W]))
    u_vec = [(U["comm_W"], r_y[1:], eval_W), (U["comm_E"], r_x, eval_E)]
    joint_u, joint_w, chal, sc_batch, evals_batch = batch_eval_reduce(p, curve, u_vec, [list(W["W"]), list(W["E"])], tr)
    return dict(sc_proof_outer=sc_outer, claims_outer=(claim_Az, claim_Bz, claim_Cz), eval_E=eval_E,
                sc_proof_inner=sc_inner, eval_W=eval_W, sc_proof_batch=sc_batch, evals_batch=evals_batch,
                r_x=r_x, r_y=r_y, batched_c=joint_u[0], batched_x=joint_u[1], batched_e=joint_u[2],
                batched_poly=joint_w, transcript=tr)


def verify_core(p, curve, S, U, vk_digest, proof):
    """snark.rs:259-396 without EE::verify.  Returns the joint (C, x, e) or raises AssertionError."""
    return _verify_core_with_transcript(p, curve, S, U, vk_digest, proof, {})


def _verify_core_with_transcript(p, curve, S, U, vk_digest, proof, holder):
    num_cons, num_vars = S["num_cons"], S["num_vars"]
    nrx, nry = _log2(num_cons), _log2(num_vars) + 1
    tr = _start_transcript(p, U, vk_digest)
    tau = [tr.squeeze(b"t") for _ in range(nrx)]
    claim_outer_final, r_x = sumcheck_verify(p, proof["sc_proof_outer"], 0, nrx, 3, tr)
    cAz, cBz, cCz = proof["claims_outer"]
    expected = eq_evaluate(p, tau, r_x) * (cAz * cBz - U["u"] * cCz - proof["eval_E"]) % p
    assert claim_outer_final == expected, "outer sum-check final claim (snark.rs:283-288)"
    tr.absorb_bytes(b"claims_outer", scalars_bytes([cAz, cBz, cCz, proof["eval_E"]]))
    r = tr.squeeze(b"r")
    claim_inner_joint = (cAz + r * cBz + r * r * cCz) % p
    claim_inner_final, r_y = sumcheck_verify(p, proof["sc_proof_inner"], claim_inner_joint, nry, 2, tr)
    eval_X = sparse_poly_evaluate(p, _log2(num_vars), [U["u"] % p] + [x % p for x in U["X"]], r_y[1:])
    eval_Z = ((1 - r_y[0]) * proof["eval_W"] + r_y[0] * eval_X) % p
    T_x, T_y = eq_evals(p, r_x), eq_evals(p, r_y)

    def mat_eval(M):  # snark.rs:327-353 multi_evaluate
        return sum(T_x[row] * T_y[col] * v for (row, col, v) in M) % p
    evA, evB, evC = (mat_eval(S[k]) for k in "ABC")
    assert claim_inner_final == (evA + r * evB + r * r * evC) * eval_Z % p, "inner sum-check final claim (snark.rs:357-360)"
    tr.absorb_bytes(b"w", scalars_bytes([proof["eval_W"]]))
    u_vec = [(U["comm_W"], r_y[1:], proof["eval_W"]), (U["comm_E"], r_x, proof["eval_E"])]
    # batch_eval_verify, spartan/mod.rs:436-480
    rho = tr.squeeze(b"r")
    powers = [pow(rho, i, p) for i in range(2)]
    num_rounds = [len(x) for (_, x, _) in u_vec]
    nmax = max(num_rounds)
    claim = sum(e * pow(2, nmax - n, p) * k for (_, _, e), n, k in zip(u_vec, num_rounds, powers)) % p  # verify_batch
    claim_batch_final, r_b = sumcheck_verify(p, proof["sc_proof_batch"], claim, nmax, 2, tr)
    exp = 0
    for (_, x, _), ev, k in zip(u_vec, proof["evals_batch"], powers):
        r_hi = r_b[nmax - len(x):]
        exp += eq_evaluate(p, r_hi, x) * ev * k
    assert claim_batch_final == exp % p, "batch evaluation final claim (spartan/mod.rs:455-470)"
    tr.absorb_bytes(b"l", scalars_bytes(proof["evals_batch"]))
    c = tr.squeeze(b"c")
    holder["tr"] = tr
    return batch_diff_size_instance(p, curve, [cm for (cm, _, _) in u_vec], proof["evals_batch"], num_rounds, r_b, c)


# ---- the whole SNARK: prove_core + EE::prove (HyperKZG), verify_core + EE::verify ---------------------------
def prove(p, curve, cid, ck_bytes, S, U, W, vk_digest):
    """RelaxedR1CSSNARK::prove (snark.rs:113-256) with the HyperKZG evaluation argument
    (hyperkzg.rs:926-1116) on the batched claim; `ck_bytes` = the SRS the commitments in U were made with."""
    from . import hyperkzg_ref as hk
    from .pyref import mont_bytes
    proof = prove_core(p, curve, S, U, W, vk_digest)
    hat_P = b"".join(mont_bytes(p, v) for v in proof["batched_poly"])
    proof["eval_arg"] = hk.prove(cid, ck_bytes, hat_P, proof["batched_x"], proof["transcript"])
    return proof


def verify(p, curve, cid, tau, S, U, vk_digest, proof) -> bool:
    """RelaxedR1CSSNARK::verify (snark.rs:259-396) incl. EE::verify, the pairing check replaced by the
    equivalent group equation for a test SRS with known tau (oracle/hyperkzg_ref.py)."""
    from . import hyperkzg_ref as hk
    from .ppsnark_ref import commitments_bytes  # noqa: F401  (same encodings as verify_core)
    tr_holder = {}

    # verify_core re-derives the transcript; run it again here to continue into EE::verify
    C, x, e = _verify_core_with_transcript(p, curve, S, U, vk_digest, proof, tr_holder)
    return hk.verify(cid, tau, C, x, e, proof["eval_arg"], tr_holder["tr"])


def prove_ipa(p, curve, ck_pts, ck_c, S, U, W, vk_digest):
    """The secondary-curve half of CompressedSNARK (S2 = snark + IPA, nova/mod.rs:872-880): prove_core, then
    EvaluationEngine::prove of provider/ipa_pc.rs:64-77 -- an inner-product argument between the batched
    polynomial and the eq table of the batched point."""
    from .pyref import ipa_prove
    proof = prove_core(p, curve, S, U, W, vk_digest)
    b_vec = eq_evals(p, proof["batched_x"])
    proof["eval_arg"] = ipa_prove(curve, ck_pts, ck_c, proof["batched_c"], b_vec, proof["batched_e"],
                                  proof["batched_poly"], proof["transcript"])
    return proof


def verify_ipa(p, curve, ck_pts, ck_c, S, U, vk_digest, proof) -> bool:
    """RelaxedR1CSSNARK::verify with the IPA evaluation engine (ipa_pc.rs:80-100, 286-396)."""
    from .pyref import ipa_verify
    holder = {}
    C, x, e = _verify_core_with_transcript(p, curve, S, U, vk_digest, proof, holder)
    L_vec, R_vec, a_hat = proof["eval_arg"]
    return ipa_verify(curve, ck_pts, ck_c, C, eq_evals(p, x), e, L_vec, R_vec, a_hat, holder["tr"])
