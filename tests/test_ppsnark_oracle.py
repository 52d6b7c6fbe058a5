"""The CPU restatement of the ppsnark prover (oracle/ppsnark_ref.py) is pinned by the restated
VERIFIER: a proof over a random satisfying relaxed R1CS instance must pass both sum-check final
checks (ppsnark.rs:1386-1600), and tampering with it must fail."""
import pytest

from oracle import ppsnark_ref as pr
from oracle.ppsnark_ref import random_instance
from oracle.pyref import CURVES, FIELD_MODULUS, SplitMix64


@pytest.mark.parametrize("fid,num_cons,num_vars", [(0, 8, 8), (3, 16, 8), (1, 4, 16)])
def test_prover_restatement_verifies(fid, num_cons, num_vars):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(100 + fid)
    S, W, u, X = random_instance(p, rng, num_cons, num_vars, num_io=2)
    spark = pr.SparkRepr(p, S["A"], S["B"], S["C"], num_cons, num_vars)
    # commitments only feed the transcript here: a cheap injective stand-in is enough for this test
    commit = lambda v: (sum((i + 1) * x for i, x in enumerate(v)) % p, len(v))
    U = dict(comm_W=commit(W["W"]), comm_E=commit(W["E"]), u=u, X=X)
    proof = pr.prove_core(p, commit, S, spark, U, W, vk_digest=12345)
    assert pr.verify_core(p, num_cons, num_vars, spark.N, U, 12345, proof)
    # batched opening claim is consistent
    from oracle.pyref import mle_evaluate
    assert mle_evaluate(p, proof["batched_poly"], proof["r_inner_batched"]) == proof["batched_eval"]
    # tampering is caught
    bad = dict(proof)
    bad["eval_L_row"] = (bad["eval_L_row"] + 1) % p
    with pytest.raises(AssertionError):
        pr.verify_core(p, num_cons, num_vars, spark.N, U, 12345, bad)


def test_unsatisfied_instance_fails():
    p = FIELD_MODULUS[0]
    rng = SplitMix64(7)
    S, W, u, X = random_instance(p, rng, 8, 8, 2)
    W["E"][3] = (W["E"][3] + 1) % p
    spark = pr.SparkRepr(p, S["A"], S["B"], S["C"], 8, 8)
    commit = lambda v: (sum((i + 1) * x for i, x in enumerate(v)) % p, len(v))
    U = dict(comm_W=commit(W["W"]), comm_E=commit(W["E"]), u=u, X=X)
    proof = pr.prove_core(p, commit, S, spark, U, W, vk_digest=1)
    with pytest.raises(AssertionError):
        pr.verify_core(p, 8, 8, spark.N, U, 1, proof)
