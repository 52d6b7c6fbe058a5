"""Shared body: the whole MicroSpartan prover incl. the HyperKZG evaluation argument over a test SRS (GPU:
tests/test_zz_new_paths_gpu.py; CPU, emulated device: tests/test_ppsnark_mirror_cpu.py): the mirror's proof equals
the oracle's and the restated verifier -- both sum-check claims, the batched opening claim over the 15
commitments and the KZG equation -- accepts it."""
from oracle import hyperkzg_ref as hk
from oracle import ppsnark_ref as pr
from oracle.pyref import CURVES, Keccak256Transcript, SplitMix64, mont_bytes
from test_ppsnark_gpu import csr


def run(nb, oracle, num_cons, num_vars, device_transcript):
    from nova_b200 import ppsnark as dp
    from nova_b200 import spartan as sp
    cid = 0
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    rng = SplitMix64(3300 + num_cons)
    S, W, u, X = pr.random_instance(p, rng, num_cons, num_vars, num_io=2)
    spark_ref = pr.SparkRepr(p, S["A"], S["B"], S["C"], num_cons, num_vars)
    N = spark_ref.N
    tau = rng.field(p)
    srs = hk.setup_srs(cid, N, tau)

    def commit_ref(v):
        return c.affine_from_bytes(oracle.msm(cid, pack(v), srs[:64 * len(v)]))
    U = dict(comm_W=commit_ref(W["W"]), comm_E=commit_ref(W["E"]), u=u, X=X)
    ref = pr.prove(p, c, cid, srs, commit_ref, S, spark_ref, U, W, 4711)
    S_comm = pr.shape_commitments(commit_ref, spark_ref)
    assert pr.verify(p, c, cid, tau, num_cons, num_vars, N, U, S_comm, 4711, ref)
    ck = nb.CommitmentKey(nb.Curve(cid), srs)
    ncols = num_vars + 1 + len(X)
    mats = {}
    for name in "ABC":
        d, idx, ptr = csr(S[name], num_cons)
        mats[name] = sp.SparseMatrix(fid, pack(d), idx, ptr, ncols)
    spark = dp.SparkRepr(fid, S["A"], S["B"], S["C"], num_cons, num_vars)
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    got = dp.prove(nb.Curve(cid), ck, dict(num_cons=num_cons, num_vars=num_vars, **mats), spark, U,
                   dict(W=pack(W["W"]), E=pack(W["E"])), 4711, tr, device_transcript=device_transcript)
    gc_, gw, gv = got["eval_arg"]
    rc_, rw, rv = ref["eval_arg"]
    assert list(gc_) == list(rc_) and list(gw) == list(rw) and [list(t) for t in gv] == [list(t) for t in rv]
    assert pr.verify(p, c, cid, tau, num_cons, num_vars, N, U, S_comm, 4711, got)
    bad = dict(got, eval_W=(got["eval_W"] + 1) % p)
    try:
        ok = pr.verify(p, c, cid, tau, num_cons, num_vars, N, U, S_comm, 4711, bad)
    except AssertionError:
        ok = False
    assert not ok
    ck.release()
