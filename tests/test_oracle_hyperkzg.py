"""The oracle's HyperKZG prover restatement must satisfy the verifier restatement
(oracle/hyperkzg_ref.py: hyperkzg.rs:1119-1242 with the pairing replaced by L = [tau]R for a test
SRS whose tau is known), and the reference's hand cases must hold (hyperkzg.rs:1265-1327)."""
import pytest

from oracle import hyperkzg_ref as hk
from oracle.pyref import CURVES, Keccak256Transcript, SplitMix64, mle_evaluate, mont_bytes


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


@pytest.mark.parametrize("ell", [2, 3, 5])
def test_prover_restatement_verifies(oracle, ell):
    cid = 0
    c = CURVES[cid]
    p = c.q
    rng = SplitMix64(300 + ell)
    n = 1 << ell
    tau = rng.field(p)
    ck = hk.setup_srs(cid, n, tau)
    poly = [rng.field(p) for _ in range(n)]
    x = [rng.field(p) for _ in range(ell)]
    y = mle_evaluate(p, poly, x)
    C = c.affine_from_bytes(oracle.msm(cid, pack(p, poly), ck))
    r, q, d0 = rng.field(p), rng.field(p), rng.field(p)
    com, v, w = hk.prove_core(cid, ck, pack(p, poly), x, r, q)
    assert hk.verify_core(cid, tau, C, x, y, com, v, w, r, q, d0)
    # every kind of message is bound by the check
    assert not hk.verify_core(cid, tau, C, x, (y + 1) % p, com, v, w, r, q, d0)
    v_bad = [list(t) for t in v]
    v_bad[0][2] = (v_bad[0][2] + 1) % p
    assert not hk.verify_core(cid, tau, C, x, y, com, v_bad, w, r, q, d0)
    assert not hk.verify_core(cid, tau, C, x, y, com, v, [w[1], w[0], w[2]], r, q, d0)
    if ell > 2:
        assert not hk.verify_core(cid, tau, C, x, y, [com[1], com[0]] + com[2:], v, w, r, q, d0)


def test_reference_hand_cases(oracle):
    """poly [1,2,1,4] at (4,3) -> 28 and poly [1,2,2,4] at the four corner-ish points
    (hyperkzg.rs:1265-1327): the evaluation the argument proves is the MLE value."""
    cid = 0
    c = CURVES[cid]
    p = c.q
    assert mle_evaluate(p, [1, 2, 1, 4], [4, 3]) == 28
    for pt, val in (([0, 0], 1), ([0, 1], 2), ([1, 1], 4), ([0, 2], 3), ([2, 2], 9)):
        assert mle_evaluate(p, [1, 2, 2, 4], pt) == val
    tau = 0x1234567
    ck = hk.setup_srs(cid, 4, tau)
    # test_hyperkzg_eval (:1265-1313): five accepting (point, eval) pairs, two rejecting ones
    poly = [1, 2, 2, 4]
    C = c.affine_from_bytes(oracle.msm(cid, pack(p, poly), ck))
    for pt, val, ok in (([0, 0], 1, True), ([0, 1], 2, True), ([1, 1], 4, True), ([0, 2], 3, True),
                        ([2, 2], 9, True), ([2, 2], 50, False), ([0, 2], 4, False)):
        com, v, w = hk.prove_core(cid, ck, pack(p, poly), pt, 1234577, 99)
        assert hk.verify_core(cid, tau, C, pt, val, com, v, w, 1234577, 99, 5) == ok, (pt, val)
    # test_hyperkzg_small (:1317-1327)
    poly, x = [1, 2, 1, 4], [4, 3]
    C = c.affine_from_bytes(oracle.msm(cid, pack(p, poly), ck))
    com, v, w = hk.prove_core(cid, ck, pack(p, poly), x, 77, 99)
    assert hk.verify_core(cid, tau, C, x, 28, com, v, w, 77, 99, 5)
    assert not hk.verify_core(cid, tau, C, x, 29, com, v, w, 77, 99, 5)


def test_reference_hand_cases_with_the_transcript(oracle):
    """test_hyperkzg_eval / test_hyperkzg_small (hyperkzg.rs:1265-1327) as the reference runs them: prover and
    verifier each start a Keccak transcript b"TestEval" and derive r, q, d_0 from it."""
    cid = 0
    c = CURVES[cid]
    p = c.q
    tau = 0xABCDEF12345
    ck = hk.setup_srs(cid, 4, tau)

    def run(poly, pt, val):
        C = c.affine_from_bytes(oracle.msm(cid, pack(p, poly), ck))
        tp, tv = Keccak256Transcript(p, b"TestEval"), Keccak256Transcript(p, b"TestEval")
        proof = hk.prove(cid, ck, pack(p, poly), pt, tp)
        ok = hk.verify(cid, tau, C, pt, val, proof, tv)
        assert tp.squeeze(b"s") == tv.squeeze(b"s")  # both sides end in the same transcript state
        return ok
    for pt, val, ok in (([0, 0], 1, True), ([0, 1], 2, True), ([1, 1], 4, True), ([0, 2], 3, True), ([2, 2], 9, True),
                        ([2, 2], 50, False), ([0, 2], 4, False)):
        assert run([1, 2, 2, 4], pt, val) == ok, (pt, val)
    assert run([1, 2, 1, 4], [4, 3], 28) and not run([1, 2, 1, 4], [4, 3], 29)


@pytest.mark.parametrize("ell", [4, 5, 6])
def test_random_polys_with_the_transcript(oracle, ell):
    """test_hyperkzg_large (hyperkzg.rs:1380-1416): random polynomial and point, prove -> verify, then a
    tampered proof (one evaluation changed) must fail."""
    cid = 0
    c = CURVES[cid]
    p = c.q
    rng = SplitMix64(ell)
    n = 1 << ell
    tau = rng.field(p)
    ck = hk.setup_srs(cid, n, tau)
    poly = [rng.field(p) for _ in range(n)]
    x = [rng.field(p) for _ in range(ell)]
    y = mle_evaluate(p, poly, x)
    C = c.affine_from_bytes(oracle.msm(cid, pack(p, poly), ck))
    proof = hk.prove(cid, ck, pack(p, poly), x, Keccak256Transcript(p, b"TestEval"))
    assert hk.verify(cid, tau, C, x, y, proof, Keccak256Transcript(p, b"TestEval"))
    com, w, v = proof
    v_bad = [list(t) for t in v]
    v_bad[1][1] = (v_bad[1][1] + 1) % p
    assert not hk.verify(cid, tau, C, x, y, (com, w, v_bad), Keccak256Transcript(p, b"TestEval"))
    assert not hk.verify(cid, tau, C, x, y, proof, Keccak256Transcript(p, b"OtherLabel"))
