"""Parity of the device ppsnark prover core (nova_b200/ppsnark.py: SpMV, outer sum-check, spark
oracles, LogUp fingerprints + batch inversion, six commitments, the three-engine batched sum-check,
final evaluations and the batched opening polynomial) against the CPU restatement
(oracle/ppsnark_ref.py), which is itself pinned by the restated verifier (test_ppsnark_oracle.py).
Every prover message, challenge, evaluation and commitment must be identical, and the device proof
must pass the restated verifier."""
import pytest

from oracle import ppsnark_ref as pr
from oracle.pyref import CURVES, FIELD_MODULUS, Keccak256Transcript, SplitMix64, from_mont_bytes, mont_bytes
from oracle.ppsnark_ref import random_instance

pytestmark = pytest.mark.gpu


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


def csr(M, rows):
    """(row, col, val) triplets in row order -> CSR arrays."""
    data, indices, indptr = [], [], [0]
    k = 0
    for r in range(rows):
        while k < len(M) and M[k][0] == r:
            data.append(M[k][2])
            indices.append(M[k][1])
            k += 1
        indptr.append(len(indices))
    return data, indices, indptr


@pytest.mark.parametrize("cid,num_cons,num_vars", [(0, 8, 8), (1, 16, 8), (3, 4, 16), (0, 64, 64), (2, 256, 128)])
def test_prove_core_matches_oracle(b200, oracle, cid, num_cons, num_vars, device_transcript=False):
    from nova_b200 import ppsnark as dp
    from nova_b200 import spartan as sp
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    rng = SplitMix64(900 + cid + num_cons)
    S, W, u, X = random_instance(p, rng, num_cons, num_vars, num_io=2)
    spark_ref = pr.SparkRepr(p, S["A"], S["B"], S["C"], num_cons, num_vars)
    N = spark_ref.N
    bases = oracle.gen_bases(cid, N)
    ck = b200.CommitmentKey(b200.Curve(cid), bases, None, 0)

    def commit_ref(v):
        return c.affine_from_bytes(oracle.msm(cid, pack(p, v), bases[:64 * len(v)]))
    U = dict(comm_W=commit_ref(W["W"]), comm_E=commit_ref(W["E"]), u=u, X=X)
    ref = pr.prove_core(p, commit_ref, S, spark_ref, U, W, vk_digest=777)

    ncols = num_vars + 1 + len(X)
    mats = {}
    for name in "ABC":
        d, idx, ptr = csr(S[name], num_cons)
        mats[name] = sp.SparseMatrix(fid, pack(p, d), idx, ptr, ncols)
    Sd = dict(num_cons=num_cons, num_vars=num_vars, **mats)
    spark = dp.SparkRepr(fid, S["A"], S["B"], S["C"], num_cons, num_vars)
    assert spark.N == N
    Wd = dict(W=pack(p, W["W"]), E=pack(p, W["E"]))
    got = dp.prove_core(b200.Curve(cid), ck, Sd, spark, U, Wd, 777, Keccak256Transcript(p, b"RelaxedR1CSSNARK"),
                        device_transcript=device_transcript)

    for k in ref:
        if k in ("batched_poly", "transcript"):
            continue
        assert got[k] == ref[k], k
    bp = got["batched_poly"].to_bytes(32 * N)
    assert [from_mont_bytes(p, bp[i:i + 32]) for i in range(0, len(bp), 32)] == ref["batched_poly"]
    assert pr.verify_core(p, num_cons, num_vars, N, U, 777, got)


def test_logup_hash_and_vec_mul(b200, oracle):
    """k_logup_hash (own-address and explicit-address forms) and k_vec_mul against integers."""
    import ctypes
    from nova_b200 import fields, ppsnark as dp
    from nova_b200.native import check, lib
    from nova_b200.spartan import DeviceVec
    for fid in (0, 3):
        p = FIELD_MODULUS[fid]
        rng = SplitMix64(31 + fid)
        n = 1000
        val = [rng.field(p) for _ in range(n)]
        addr = [rng.next() % n for _ in range(n)]
        g, r = rng.field(p), rng.field(p)
        dv, da = DeviceVec.from_bytes(pack(p, val)), DeviceVec.from_bytes(pack(p, addr))
        dg, dr = dp.dev_scalar(fid, g), dp.dev_scalar(fid, r)
        out = DeviceVec(32 * n)
        check(lib().b200_logup_hash_dev(fid, dv.ptr, None, dg.ptr, dr.ptr, n, out.ptr, None))
        assert fields.unpack(fid, out.to_bytes()) == [(val[i] * g + i + r) % p for i in range(n)]
        check(lib().b200_logup_hash_dev(fid, dv.ptr, da.ptr, dg.ptr, dr.ptr, n, out.ptr, None))
        assert fields.unpack(fid, out.to_bytes()) == [(val[i] * g + addr[i] + r) % p for i in range(n)]
        check(lib().b200_vec_mul_dev(fid, dv.ptr, da.ptr, n, out.ptr, None))
        assert fields.unpack(fid, out.to_bytes()) == [val[i] * addr[i] % p for i in range(n)]
