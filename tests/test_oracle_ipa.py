"""The oracle's literal IPA prover (pyref.ipa_prove, ipa_pc.rs:174-285) is pinned by a restatement
of the reference VERIFIER (pyref.ipa_verify, ipa_pc.rs:286-396): honest proofs pass, altered ones
fail.  tests/test_ipa_gpu.py then compares the key-folding-free CUDA prover with ipa_prove."""
import pytest

from oracle.pyref import CURVES, Keccak256Transcript, SplitMix64, ipa_prove, ipa_verify


@pytest.mark.parametrize("cid,l", [(1, 1), (1, 3), (3, 4), (2, 2)])
def test_ipa_prover_restatement_verifies(cid, l):
    c = CURVES[cid]
    q = c.q
    n = 1 << l
    rng = SplitMix64(600 + cid + l)
    ck = c.bases_arith(n + 1)
    ck_pts, ck_c = ck[:n], ck[n]
    a = [rng.field(q) for _ in range(n)]
    b = [rng.field(q) for _ in range(n)]
    claim = sum(x * y for x, y in zip(a, b)) % q
    comm_a = c.msm_naive(a, ck_pts)
    L, R, a_hat = ipa_prove(c, ck_pts, ck_c, comm_a, b, claim, a, Keccak256Transcript(q, b"ipa"))
    assert ipa_verify(c, ck_pts, ck_c, comm_a, b, claim, L, R, a_hat, Keccak256Transcript(q, b"ipa"))
    assert not ipa_verify(c, ck_pts, ck_c, comm_a, b, (claim + 1) % q, L, R, a_hat, Keccak256Transcript(q, b"ipa"))
    assert not ipa_verify(c, ck_pts, ck_c, comm_a, b, claim, L, R, (a_hat + 1) % q, Keccak256Transcript(q, b"ipa"))
    if l > 1:
        assert not ipa_verify(c, ck_pts, ck_c, comm_a, b, claim, [L[1], L[0]] + L[2:], R, a_hat,
                              Keccak256Transcript(q, b"ipa"))
