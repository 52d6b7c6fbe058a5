"""Every sum-check prover restated in oracle/pyref.py is pinned by the reference VERIFIER's
equations (SumcheckProof::verify / verify_batch, sumcheck.rs:87-160, and the final-claim checks of
the callers): honest transcripts verify, a perturbed round polynomial or final evaluation does not.

  prove_quad_prod                  sum A*B             final: A(r) B(r)              (sumcheck.rs:199-242)
  prove_cubic_with_three_inputs    sum eq(tau)(AB - C) final: eq(tau, r)(A B - C)(r) (sumcheck.rs:446-507,
                                                       checked as in ppsnark.rs:1408-1418)
  prove_batch_eval                 sum_i rho^i 2^(n-n_i) e_i, final: sum_i rho^i eq(r_hi, x_i) P_i(r_hi)
                                                       (sumcheck.rs:251-351, spartan/mod.rs:436-473)
"""
import pytest

from oracle.ppsnark_ref import eq_evaluate, sumcheck_verify
from oracle.pyref import (FIELD_MODULUS, Keccak256Transcript, SplitMix64, eq_evals, mle_evaluate, prove_batch_eval,
                          prove_batched_cubic, prove_cubic_with_three_inputs, prove_quad_prod)


@pytest.mark.parametrize("fid,ell", [(0, 1), (0, 4), (3, 6)])
def test_quad_prod(fid, ell):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(fid + ell)
    n = 1 << ell
    A = [rng.field(p) for _ in range(n)]
    B = [rng.field(p) for _ in range(n)]
    claim = sum(a * b for a, b in zip(A, B)) % p
    polys, rs, finals = prove_quad_prod(p, claim, ell, A, B, Keccak256Transcript(p, b"t"))
    e, rv = sumcheck_verify(p, polys, claim, ell, 2, Keccak256Transcript(p, b"t"))
    assert rv == rs
    assert finals == [mle_evaluate(p, A, rs), mle_evaluate(p, B, rs)]
    assert e == finals[0] * finals[1] % p
    e_bad, _ = sumcheck_verify(p, polys, (claim + 1) % p, ell, 2, Keccak256Transcript(p, b"t"))
    assert e_bad != finals[0] * finals[1] % p


@pytest.mark.parametrize("fid,ell", [(0, 2), (1, 5), (2, 7)])
def test_cubic_with_three_inputs(fid, ell):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(10 * fid + ell)
    n = 1 << ell
    A = [rng.field(p) for _ in range(n)]
    B = [rng.field(p) for _ in range(n)]
    C = [a * b % p for a, b in zip(A, B)]  # satisfied: claim 0 (snark.rs / ppsnark.rs outer sum-check)
    tau = [rng.field(p) for _ in range(ell)]
    polys, rs, finals = prove_cubic_with_three_inputs(p, 0, tau, A, B, C, Keccak256Transcript(p, b"t"))
    e, rv = sumcheck_verify(p, polys, 0, ell, 3, Keccak256Transcript(p, b"t"))
    assert rv == rs and finals == [mle_evaluate(p, V, rs) for V in (A, B, C)]
    assert e == eq_evaluate(p, tau, rs) * (finals[0] * finals[1] - finals[2]) % p
    # a non-zero claim: sum eq(tau, x) (A B - C')(x) for an unsatisfied C'
    C2 = list(C)
    C2[1] = (C2[1] + 5) % p
    claim = sum(w * (a * b - c) for w, a, b, c in zip(eq_evals(p, tau), A, B, C2)) % p
    polys, rs, finals = prove_cubic_with_three_inputs(p, claim, tau, A, B, C2, Keccak256Transcript(p, b"t"))
    e, _ = sumcheck_verify(p, polys, claim, ell, 3, Keccak256Transcript(p, b"t"))
    assert e == eq_evaluate(p, tau, rs) * (finals[0] * finals[1] - finals[2]) % p
    e0, _ = sumcheck_verify(p, polys, 0, ell, 3, Keccak256Transcript(p, b"t"))  # wrong claim
    assert e0 != eq_evaluate(p, tau, rs) * (finals[0] * finals[1] - finals[2]) % p


@pytest.mark.parametrize("fid,sizes", [(0, [3]), (0, [4, 2, 4]), (3, [1, 5, 3, 5])])
def test_batch_eval(fid, sizes):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(7 * fid + sum(sizes))
    polys = [[rng.field(p) for _ in range(1 << s)] for s in sizes]
    points = [[rng.field(p) for _ in range(s)] for s in sizes]
    claims = [mle_evaluate(p, P, x) for P, x in zip(polys, points)]
    rho = rng.field(p)
    coeffs = [pow(rho, i, p) for i in range(len(sizes))]
    out, rs, finals = prove_batch_eval(p, claims, sizes, polys, points, coeffs, Keccak256Transcript(p, b"t"))
    nmax = max(sizes)
    claim = sum(c * pow(2, nmax - s, p) * k for c, s, k in zip(claims, sizes, coeffs)) % p  # verify_batch
    e, rv = sumcheck_verify(p, out, claim, nmax, 2, Keccak256Transcript(p, b"t"))
    assert rv == rs
    expected = 0
    for P, x, s, k, f in zip(polys, points, sizes, coeffs, finals):
        r_hi = rs[nmax - s:]
        assert f == mle_evaluate(p, P, r_hi)
        expected += eq_evaluate(p, r_hi, x) * f * k
    assert e == expected % p
    bad = list(claims)
    bad[0] = (bad[0] + 1) % p
    claim_bad = sum(c * pow(2, nmax - s, p) * k for c, s, k in zip(bad, sizes, coeffs)) % p
    e_bad, _ = sumcheck_verify(p, out, claim_bad, nmax, 2, Keccak256Transcript(p, b"t"))
    assert e_bad != expected % p


@pytest.mark.parametrize("fid,k,ell,zero_tau", [(0, 1, 3, ()), (0, 3, 5, ()), (3, 2, 4, (0, 2))])
def test_batched_cubic(fid, k, ell, zero_tau):
    """prove_batched_cubic (sumcheck.rs:513-577): the verifier's final claim is
    eq(tau, r) * sum_i alpha_i (A_i(r) B_i(r) - C_i(r)); with K = 1 and alpha = 1 it must coincide with
    prove_cubic_with_three_inputs message by message."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(31 * fid + 7 * k + ell)
    n = 1 << ell
    As, Bs, Cs = ([[rng.field(p) for _ in range(n)] for _ in range(k)] for _ in range(3))
    alphas = [rng.field(p) for _ in range(k)]
    tau = [0 if i in zero_tau else rng.field(p) for i in range(ell)]
    w = eq_evals(p, tau)
    claim = sum(w[x] * sum(al * (A[x] * B[x] - C[x]) for al, A, B, C in zip(alphas, As, Bs, Cs)) for x in range(n)) % p
    polys, rs, finals = prove_batched_cubic(p, claim, tau, As, Bs, Cs, alphas, Keccak256Transcript(p, b"t"))
    e, rv = sumcheck_verify(p, polys, claim, ell, 3, Keccak256Transcript(p, b"t"))
    assert rv == rs
    for i in range(k):
        assert finals[i] == [mle_evaluate(p, V[i], rs) for V in (As, Bs, Cs)]
    assert e == eq_evaluate(p, tau, rs) * sum(al * (f[0] * f[1] - f[2]) for al, f in zip(alphas, finals)) % p
    e_bad, _ = sumcheck_verify(p, polys, (claim + 1) % p, ell, 3, Keccak256Transcript(p, b"t"))
    assert e_bad != e
    if k == 1:
        one = prove_batched_cubic(p, claim, tau, As, Bs, Cs, [1], Keccak256Transcript(p, b"t"))
        ref = prove_cubic_with_three_inputs(p, claim, tau, As[0], Bs[0], Cs[0], Keccak256Transcript(p, b"t"))
        assert (one[0], one[1], one[2][0]) == ref
