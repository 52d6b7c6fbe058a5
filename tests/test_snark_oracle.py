"""oracle/snark_ref.py (restatement of spartan::snark::RelaxedR1CSSNARK::prove, snark.rs:113-256) pinned
by the restated VERIFIER (snark.rs:259-396, spartan/mod.rs:436-480): honest proofs of satisfying relaxed
instances pass and reduce to a true opening claim; unsatisfied witnesses and altered messages fail."""
import copy

import pytest

from oracle.ppsnark_ref import random_instance
from oracle.pyref import CURVES, SplitMix64, mle_evaluate
from oracle.snark_ref import prove_core, verify_core


def make_case(cid, num_cons, num_vars, num_io, seed):
    c = CURVES[cid]
    p = c.q
    rng = SplitMix64(seed)
    S, W, u, X = random_instance(p, rng, num_cons, num_vars, num_io)
    bases = c.bases_arith(max(num_cons, num_vars))
    commit = lambda v: c.msm_naive(v, bases[:len(v)])
    U = dict(comm_W=commit(W["W"]), comm_E=commit(W["E"]), u=u, X=X)
    return c, p, S, U, W, commit


@pytest.mark.parametrize("cid,num_cons,num_vars,num_io", [(0, 4, 4, 1), (0, 16, 8, 2), (1, 8, 16, 2), (3, 32, 32, 3)])
def test_honest_proof_verifies_and_reduces_to_a_true_claim(cid, num_cons, num_vars, num_io):
    c, p, S, U, W, commit = make_case(cid, num_cons, num_vars, num_io, 100 + num_cons)
    proof = prove_core(p, c, S, U, W, vk_digest=12345)
    C, x, e = verify_core(p, c, S, U, 12345, proof)
    assert (C, x, e) == (proof["batched_c"], proof["batched_x"], proof["batched_e"])
    # the claim handed to EE::prove is true, and C is the commitment of the batched polynomial
    assert mle_evaluate(p, proof["batched_poly"], x) == e
    assert commit(proof["batched_poly"]) == C
    # the transcripts of prover and verifier continue identically
    # (EE::prove / EE::verify start from the same state)
    assert proof["transcript"].squeeze(b"x") is not None


def test_unsatisfied_witness_is_rejected():
    c, p, S, U, W, commit = make_case(0, 8, 8, 2, 7)
    W2 = copy.deepcopy(W)
    W2["E"][3] = (W2["E"][3] + 1) % p
    U2 = dict(U, comm_E=commit(W2["E"]))
    proof = prove_core(p, c, S, U2, W2, 1)
    with pytest.raises(AssertionError, match="outer sum-check"):
        verify_core(p, c, S, U2, 1, proof)


@pytest.mark.parametrize("field,match", [("eval_W", "inner sum-check"), ("eval_E", "outer sum-check"),
                                         ("claims_outer", "outer sum-check"), ("evals_batch", "batch evaluation"),
                                         ("sc_proof_inner", "inner sum-check"), ("sc_proof_batch", "batch evaluation")])
def test_tampered_messages_are_rejected(field, match):
    c, p, S, U, W, _ = make_case(0, 8, 8, 2, 9)
    proof = prove_core(p, c, S, U, W, 5)
    verify_core(p, c, S, U, 5, proof)
    bad = copy.deepcopy({k: v for k, v in proof.items() if k != "transcript"})
    if field == "claims_outer":
        a, b, cc = bad["claims_outer"]
        bad["claims_outer"] = (a, (b + 1) % p, cc)
    elif field == "evals_batch":
        bad["evals_batch"][1] = (bad["evals_batch"][1] + 1) % p
    elif field.startswith("sc_proof"):
        bad[field][-1][0] = (bad[field][-1][0] + 1) % p
    else:
        bad[field] = (bad[field] + 1) % p
    with pytest.raises(AssertionError, match=match):
        verify_core(p, c, S, U, 5, bad)
    with pytest.raises(AssertionError):
        verify_core(p, c, S, U, 6, proof)  # another vk digest: every challenge changes


@pytest.mark.parametrize("num_cons,num_vars", [(8, 8), (4, 16)])
def test_full_snark_with_hyperkzg_evaluation_argument(oracle, num_cons, num_vars):
    """prove = prove_core + EE::prove, verify = verify_core + EE::verify over a test SRS [tau^i]G: the honest
    proof verifies, a proof for another instance or with an altered evaluation argument does not."""
    from oracle import hyperkzg_ref as hk
    from oracle.pyref import mont_bytes
    from oracle.snark_ref import prove, verify
    cid = 0
    c = CURVES[cid]
    p = c.q
    rng = SplitMix64(77 + num_cons)
    S, W, u, X = random_instance(p, rng, num_cons, num_vars, 2)
    tau = rng.field(p)
    n_key = max(num_cons, num_vars)
    ck = hk.setup_srs(cid, n_key, tau)
    pack = lambda v: b"".join(mont_bytes(p, x) for x in v)
    commit = lambda v: c.affine_from_bytes(oracle.msm(cid, pack(v), ck[:64 * len(v)]))
    U = dict(comm_W=commit(W["W"]), comm_E=commit(W["E"]), u=u, X=X)
    proof = prove(p, c, cid, ck, S, U, W, 31337)
    assert verify(p, c, cid, tau, S, U, 31337, proof)
    com, w, v = proof["eval_arg"]
    bad = dict(proof, eval_arg=(com, [w[1], w[0], w[2]], v))
    assert not verify(p, c, cid, tau, S, U, 31337, bad)
    v2 = [list(t) for t in v]
    v2[0][0] = (v2[0][0] + 1) % p
    assert not verify(p, c, cid, tau, S, U, 31337, dict(proof, eval_arg=(com, w, v2)))
