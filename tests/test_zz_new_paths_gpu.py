"""GPU parity of the entry points added after this round's GPU budget was spent.

Their device code was validated on the CPU through the host build of the same headers
(tests/test_host_transcript.py, tests/test_host_field.py) but had not run on a B200 when it was
committed, so this file sorts LAST: with `pytest -x` a failure here cannot hide the rest of the suite.

* sum-check round loops with the Keccak transcript on the device (b200_sumcheck_quad_prod,
  b200_sumcheck_cubic3, b200_sc_round_dev): every prover message, challenge, final evaluation and the
  transcript state afterwards must equal the oracle's (sumcheck.rs:199-242, 446-507).
"""
import pytest

from oracle.pyref import (FIELD_MODULUS, Keccak256Transcript, SplitMix64, eq_evals, mont_bytes,
                          prove_cubic_with_three_inputs, prove_quad_prod)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def sp(b200):
    from nova_b200 import spartan
    return spartan


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


def _transcripts(p, label, absorbs, rng):
    a, b = Keccak256Transcript(p, label), Keccak256Transcript(p, label)
    for _ in range(absorbs):
        x = rng.field(p)
        a.absorb_scalar(b"x", x)
        b.absorb_scalar(b"x", x)
    return a, b


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
@pytest.mark.parametrize("l,absorbs", [(1, 0), (5, 2), (11, 0), (14, 40)])
def test_quad_prod_device_transcript(sp, fid, l, absorbs):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(500 + 7 * l + fid)
    A = [rng.field(p) for _ in range(1 << l)]
    B = [rng.field(p) for _ in range(1 << l)]
    claim = sum(a * b for a, b in zip(A, B)) % p
    t_ref, t_dev = _transcripts(p, b"dq", absorbs, rng)
    exp = prove_quad_prod(p, claim, l, A, B, t_ref)
    got = sp.SumcheckProof.prove_quad_prod_device(fid, claim, l, pack(p, A), pack(p, B), t_dev)
    assert got[1] == exp[1], "challenges"
    assert got[0] == exp[0], "compressed round polynomials"
    assert got[2] == exp[2], "final evaluations"
    assert t_dev.squeeze(b"n") == t_ref.squeeze(b"n"), "transcript state after the loop"


@pytest.mark.parametrize("fid", [0, 3])
@pytest.mark.parametrize("l,zero_tau_at", [(1, None), (2, None), (3, 1), (6, None), (6, 0), (7, 6), (12, None), (13, 4)])
def test_cubic3_device_transcript(sp, fid, l, zero_tau_at):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(600 + 5 * l + fid)
    n = 1 << l
    A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
    taus = [rng.field(p) for _ in range(l)]
    if zero_tau_at is not None:
        taus[zero_tau_at] = 0
    eqt = eq_evals(p, taus)
    claim = sum(e * (a * b - c) for e, a, b, c in zip(eqt, A, B, C)) % p
    t_ref, t_dev = _transcripts(p, b"dc", 1, rng)
    exp = prove_cubic_with_three_inputs(p, claim, taus, A, B, C, t_ref)
    got = sp.SumcheckProof.prove_cubic_with_three_inputs_device(fid, claim, taus, pack(p, A), pack(p, B), pack(p, C),
                                                                t_dev)
    assert got[1] == exp[1], "challenges"
    assert got[0] == exp[0], "compressed round polynomials"
    assert got[2] == exp[2], "final evaluations"
    assert t_dev.squeeze(b"n") == t_ref.squeeze(b"n")


def test_device_loop_equals_host_loop_large(sp):
    """2^18 elements: the device-transcript loop against the existing host-transcript loop (itself
    parity-tested against the oracle at small sizes)."""
    fid, l = 0, 18
    p = FIELD_MODULUS[fid]
    from oracle import coracle as co
    A, B, C = (co.gen_scalars(fid, 70 + k, 1 << l) for k in range(3))
    rng = SplitMix64(77)
    taus = [rng.field(p) for _ in range(l)]
    claim = rng.field(p)  # the prover does not check the claim
    t1, t2 = Keccak256Transcript(p, b"big"), Keccak256Transcript(p, b"big")
    exp = sp.SumcheckProof.prove_cubic_with_three_inputs(fid, claim, taus, A, B, C, t1)
    got = sp.SumcheckProof.prove_cubic_with_three_inputs_device(fid, claim, taus, A, B, C, t2)
    assert got == exp
    assert t1.squeeze(b"n") == t2.squeeze(b"n")
    t1, t2 = Keccak256Transcript(p, b"big"), Keccak256Transcript(p, b"big")
    exp = sp.SumcheckProof.prove_quad_prod(fid, claim, l, A, B, t1)
    got = sp.SumcheckProof.prove_quad_prod_device(fid, claim, l, A, B, t2)
    assert got == exp


def test_pending_limit_is_an_error(sp):
    from nova_b200.native import B200Error
    fid = 0
    p = FIELD_MODULUS[fid]
    t = Keccak256Transcript(p, b"x")
    t.absorb_bytes(b"big", bytes(4000))
    with pytest.raises(B200Error):
        sp.SumcheckProof.prove_quad_prod_device(fid, 1, 1, pack(p, [1, 2]), pack(p, [3, 4]), t)
