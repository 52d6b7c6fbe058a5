"""Inner-product argument prover: the device prover that never folds the key
emits exactly the L_vec / R_vec / a_hat of the literal restatement of ipa_pc.rs:174-285 (which folds
the key each round with Python big-integer group arithmetic)."""
import pytest

from oracle.pyref import CURVES, Keccak256Transcript, SplitMix64, ipa_prove, mont_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cid,l", [(1, 1), (1, 5), (3, 4), (0, 3)])
def test_ipa_prove_matches_restatement(b200, oracle, cid, l):
    from nova_b200.ipa import InnerProductArgument
    c = CURVES[cid]
    q = c.q
    n = 1 << l
    rng = SplitMix64(31 * cid + l)
    bases = c.bases_arith(n + 1, k0=777)
    ck_pts, ck_c = bases[:n], bases[n]
    a = [rng.field(q) for _ in range(n)]
    b = [rng.field(q) for _ in range(n)]
    a[0] = 0  # a zero and a one among the witness entries
    a[-1] = 1
    claim = sum(x * y for x, y in zip(a, b)) % q
    comm_a = c.msm_naive(a, ck_pts)
    exp = ipa_prove(c, ck_pts, ck_c, comm_a, b, claim, a, Keccak256Transcript(q, b"ipa"))
    key = b200.CommitmentKey(b200.Curve(cid), b"".join(c.affine_bytes(P) for P in ck_pts), c.affine_bytes(ck_c))
    pk = lambda xs: b"".join(mont_bytes(q, x) for x in xs)
    got = InnerProductArgument.prove(cid, key, comm_a, pk(b), claim, pk(a), Keccak256Transcript(q, b"ipa"))
    assert got == exp
    # the device-side inner product itself (form 11) against the oracle
    import ctypes
    from nova_b200.native import check, lib
    out = ctypes.create_string_buffer(96)
    buf = lambda x: ctypes.create_string_buffer(x, len(x))
    check(lib().b200_sc_eval(c.scalar_field, 11, buf(pk(a)), buf(pk(b)), None, n, None, 0, None, 0, 0, out))
    assert out.raw[:32] == mont_bytes(q, claim)
