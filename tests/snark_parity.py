"""Shared body of the snark prover-core parity check (GPU: tests/test_zz_new_paths_gpu.py; CPU with the
emulated device: tests/test_snark_mirror_cpu.py): every message of nova_b200.snark.prove_core equals
oracle/snark_ref.py (pinned by the restated verifier, tests/test_snark_oracle.py), the proof passes that
verifier, and the batched opening polynomial evaluates to the joint claim."""
from oracle import snark_ref as sr
from oracle.ppsnark_ref import random_instance
from oracle.pyref import CURVES, Keccak256Transcript, SplitMix64, from_mont_bytes, mle_evaluate, mont_bytes


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


def csr(M, rows):
    data, indices, indptr, k = [], [], [0], 0
    for r in range(rows):
        while k < len(M) and M[k][0] == r:
            data.append(M[k][2])
            indices.append(M[k][1])
            k += 1
        indptr.append(len(indices))
    return data, indices, indptr


def run_case(nb, oracle, cid, num_cons, num_vars, num_io, device_transcript):
    from nova_b200 import snark as ds
    from nova_b200 import spartan as sp
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    rng = SplitMix64(1300 + cid + num_cons)
    S, W, u, X = random_instance(p, rng, num_cons, num_vars, num_io)
    bases = oracle.gen_bases(cid, max(num_cons, num_vars))

    def commit_ref(v):
        return c.affine_from_bytes(oracle.msm(cid, pack(p, v), bases[:64 * len(v)]))
    U = dict(comm_W=commit_ref(W["W"]), comm_E=commit_ref(W["E"]), u=u, X=X)
    ref = sr.prove_core(p, c, S, U, W, vk_digest=4242)
    ncols = num_vars + 1 + num_io
    mats = {}
    for name in "ABC":
        d, idx, ptr = csr(S[name], num_cons)
        mats[name] = sp.SparseMatrix(fid, pack(p, d), idx, ptr, ncols)
    Sd = dict(num_cons=num_cons, num_vars=num_vars, **mats)
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    got = ds.prove_core(nb.Curve(cid), None, Sd, U, dict(W=pack(p, W["W"]), E=pack(p, W["E"])), 4242, tr,
                        device_transcript=device_transcript)
    for k in ("sc_proof_outer", "sc_proof_inner", "sc_proof_batch", "evals_batch", "r_x", "r_y", "batched_x"):
        assert [list(q) if isinstance(q, (list, tuple)) else q for q in got[k]] == \
               [list(q) if isinstance(q, (list, tuple)) else q for q in ref[k]], k
    assert tuple(got["claims_outer"]) == tuple(ref["claims_outer"])
    for k in ("eval_E", "eval_W", "batched_e"):
        assert got[k] == ref[k], k
    assert got["batched_c"] == ref["batched_c"]
    n = max(num_cons, num_vars)
    bp = got["batched_poly"].to_bytes(32 * n)
    poly = [from_mont_bytes(p, bp[i:i + 32]) for i in range(0, len(bp), 32)]
    assert poly == ref["batched_poly"]
    assert mle_evaluate(p, poly, got["batched_x"]) == got["batched_e"]
    assert sr.verify_core(p, c, S, U, 4242, got) == (got["batched_c"], got["batched_x"], got["batched_e"])
    assert tr.squeeze(b"x") == ref["transcript"].squeeze(b"x")  # EE::prove would start from the same state


def run_full(nb, oracle, num_cons, num_vars, num_io, device_transcript):
    """The whole RelaxedR1CSSNARK::prove on BN254 with the HyperKZG evaluation argument over a test SRS
    [tau^i] G: the mirror's proof (incl. eval_arg) equals the oracle's and the restated verifier -- sum-check
    claims, batch claim and the KZG opening equation -- accepts it."""
    from nova_b200 import snark as ds
    from nova_b200 import spartan as sp
    from oracle import hyperkzg_ref as hk
    cid = 0
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    rng = SplitMix64(2100 + num_cons + num_vars)
    S, W, u, X = random_instance(p, rng, num_cons, num_vars, num_io)
    tau = rng.field(p)
    n_key = max(num_cons, num_vars)
    srs = hk.setup_srs(cid, n_key, tau)

    def commit_ref(v):
        return c.affine_from_bytes(oracle.msm(cid, pack(p, v), srs[:64 * len(v)]))
    U = dict(comm_W=commit_ref(W["W"]), comm_E=commit_ref(W["E"]), u=u, X=X)
    ref = sr.prove(p, c, cid, srs, S, U, W, 555)
    ncols = num_vars + 1 + num_io
    mats = {}
    for name in "ABC":
        d, idx, ptr = csr(S[name], num_cons)
        mats[name] = sp.SparseMatrix(fid, pack(p, d), idx, ptr, ncols)
    ck = nb.CommitmentKey(nb.Curve(cid), srs)
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    got = ds.prove(nb.Curve(cid), ck, dict(num_cons=num_cons, num_vars=num_vars, **mats), U,
                   dict(W=pack(p, W["W"]), E=pack(p, W["E"])), 555, tr, device_transcript=device_transcript)
    com, w, v = got["eval_arg"]
    rcom, rw, rv = ref["eval_arg"]
    assert list(com) == list(rcom) and list(w) == list(rw) and [list(t) for t in v] == [list(t) for t in rv]
    assert sr.verify(p, c, cid, tau, S, U, 555, got)
    assert tr.squeeze(b"x") == ref["transcript"].squeeze(b"x")
    ck.release()


def run_full_ipa(nb, oracle, cid, num_cons, num_vars, num_io, device_transcript):
    """The secondary-curve SNARK (S2 = spartan::snark + IPA, nova/mod.rs:872-880) over a Pedersen key: the
    mirror's proof equals the oracle's and the restated verifier incl. the IPA check (ipa_pc.rs:286-396) accepts."""
    from nova_b200 import snark as ds
    from nova_b200 import spartan as sp
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    rng = SplitMix64(2300 + cid + num_cons)
    S, W, u, X = random_instance(p, rng, num_cons, num_vars, num_io)
    n_key = max(num_cons, num_vars)
    pts = c.bases_arith(n_key + 1, k0=4242)
    ck_pts, ck_c = pts[:n_key], pts[n_key]

    def commit_ref(v):
        return c.msm_naive(v, ck_pts[:len(v)])
    U = dict(comm_W=commit_ref(W["W"]), comm_E=commit_ref(W["E"]), u=u, X=X)
    ref = sr.prove_ipa(p, c, ck_pts, ck_c, S, U, W, 808)
    assert sr.verify_ipa(p, c, ck_pts, ck_c, S, U, 808, ref)
    ncols = num_vars + 1 + num_io
    mats = {}
    for name in "ABC":
        d, idx, ptr = csr(S[name], num_cons)
        mats[name] = sp.SparseMatrix(fid, pack(p, d), idx, ptr, ncols)
    ck = nb.CommitmentKey(nb.Curve(cid), b"".join(c.affine_bytes(P) for P in ck_pts), c.affine_bytes(ck_c))
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    got = ds.prove(nb.Curve(cid), ck, dict(num_cons=num_cons, num_vars=num_vars, **mats), U,
                   dict(W=pack(p, W["W"]), E=pack(p, W["E"])), 808, tr, device_transcript=device_transcript, ee="ipa")
    assert tuple(got["eval_arg"]) == tuple(ref["eval_arg"])
    assert sr.verify_ipa(p, c, ck_pts, ck_c, S, U, 808, got)
    bad = dict(got, eval_arg=(got["eval_arg"][0], got["eval_arg"][1], (got["eval_arg"][2] + 1) % p))
    assert not sr.verify_ipa(p, c, ck_pts, ck_c, S, U, 808, bad)
    ck.release()
