#!/usr/bin/env python3
"""The prover workloads of BASELINE.json configs[2..4] at benchmark scale, each TIMED on the GPU path and then
CHECKED outside the timed region (bench.py --workload {prove_step, hyperkzg, ppsnark}; also run by the default
bench at N = 1).  The check is what makes the number count:

  hyperkzg    HyperKZG EvaluationEngine::prove (hyperkzg.rs:926-1116) on a 2^LOG2N polynomial over a test SRS
              ck[i] = [tau^i] G (hyperkzg.rs:357-376).  The proof must be ACCEPTED by the restated verifier
              (oracle/hyperkzg_ref.verify: hyperkzg.rs:1119-1242 with the pairing replaced by L = [tau] R), for the
              commitment C = commit(P) and the value y = P(x) computed by the C oracle -- nothing the GPU computed
              enters the check except the proof itself -- and a tampered proof must be rejected.  With N > 1 GPUs
              the polynomial is split by index range (sharding.sharded_hyperkzg_prove over NCCL) and the sharded
              proof must equal the single-GPU proof message for message.
  ppsnark     RelaxedR1CSSNARK::prove of spartan/ppsnark.rs:1056-1385 incl. EE::prove, sha256-like synthetic
              shape.  Accepted by the restated verifier (oracle/ppsnark_ref.verify: both sum-check final claims,
              ppsnark.rs:1386-1600, the batched opening over the 15 commitments and the KZG equation), with the
              instance and shape commitments computed by the C oracle.
  prove_step  kernel-sequence replay of RecursiveSNARK::prove_step (nova/mod.rs:456-564; tools/prove_step_replay):
              every commitment and both folded vectors must equal the C oracle's on the same inputs.

Only the check and the cpu legs touch oracle/ (as the checker / the timed CPU implementation); the timed GPU region
runs on nova_b200 alone."""
import ctypes
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np

TAU = 0x1F3A5C7E9B2D4F60718293A4B5C6D7E8F9012345_6789ABCDEF0123456789ABCD  # the test SRS's trapdoor (< r)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _max_over_ranks(seconds: float) -> float:
    d = _dist()
    if d is None:
        return seconds
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device="cuda")
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def _barrier():
    from nova_b200.native import check, lib
    check(lib().b200_sync())
    d = _dist()
    if d is not None:
        d.barrier()


# ----------------------------------------------------------------------------------------------------------
# HyperKZG (configs[3])
# ----------------------------------------------------------------------------------------------------------
def hyperkzg(log2n=22, steps=3, warmup=1, check_parity=True):
    import nova_b200 as nb
    from hyperkzg_replay import synth_poly
    from nova_b200 import fields, sharding as sh, spartan as sp
    from nova_b200.native import check, lib
    from nova_b200.transcript import Keccak256Transcript
    d = _dist()
    rank, world = (d.get_rank(), d.get_world_size()) if d else (0, 1)
    L = lib()
    curve = nb.Curve(0)
    fid = curve.scalar_field
    p = fields.MODULUS[fid]
    ell, n = log2n, 1 << log2n
    t0 = time.time()
    ck = nb.CommitmentKey.setup_tau(curve, n, TAU)  # the full key on every GPU (sharding.sharded_hyperkzg_prove)
    poly = synth_poly(n, 4)
    rng = np.random.default_rng(9)
    x = [int.from_bytes(rng.bytes(31), "little") for _ in range(ell)]
    lo, hi = rank * (n // world), (rank + 1) * (n // world)
    host_local = np.ascontiguousarray(poly[lo:hi])
    pinned = ctypes.c_void_p()
    check(L.b200_host_alloc(32 * (hi - lo), ctypes.byref(pinned)))
    ctypes.memmove(pinned, host_local.ctypes.data, 32 * (hi - lo))
    P_local = sp.DeviceVec(32 * (hi - lo))
    comm = (sh.NcclComm() if world > 1 else None)
    setup_s = time.time() - t0

    def prove_once(upload: bool, timings=None):
        tr = Keccak256Transcript(p, b"HyperKZG bench")
        if upload:  # e2e: the polynomial arrives from pinned host memory inside the timed region
            check(L.b200_memcpy_h2d(P_local.ptr, pinned, 32 * (hi - lo)))
        if world == 1:
            return sp.hyperkzg_prove(curve, ck, P_local, x, tr, timings)

        def r_of(com):
            tr.absorb_bytes(b"c", b"".join(sp._commitment_bytes(C) for C in com))
            return tr.squeeze(b"c")

        def q_of(v):
            tr.absorb_bytes(b"v", b"".join(int(e).to_bytes(32, "little") for row in v for e in row))
            return tr.squeeze(b"r")

        def after_w(w):
            tr.absorb_bytes(b"W", b"".join(sp._commitment_bytes(C) for C in w))
            tr.squeeze(b"d")
        com, v, w = sh.sharded_hyperkzg_prove(curve, ck, P_local, x, r_of, q_of, comm, after_w)
        return com, w, v

    def timed(upload: bool):
        for _ in range(warmup):
            prove_once(upload)
        best, tot, proof, phases = None, 0.0, None, None
        for _ in range(steps):
            tm = {} if world == 1 else None
            _barrier()
            t1 = time.perf_counter()
            proof = prove_once(upload, tm)
            check(L.b200_sync())
            dt = _max_over_ranks(time.perf_counter() - t1)
            tot += dt
            if best is None or dt < best:
                best, phases = dt, tm
        return tot / steps, best, proof, phases

    check(L.b200_memcpy_h2d(P_local.ptr, pinned, 32 * (hi - lo)))
    ms_dev, best_dev, proof, phases = timed(upload=False)
    ms_e2e, _, proof_e2e, _ = timed(upload=True)
    out = {"workload": f"HyperKZG EvaluationEngine::prove, BN254, 2^{log2n} uniform scalars, test SRS [tau^i]G, "
                       f"{world} GPU(s)" + (", polynomial by index range, NCCL" if world > 1 else ""),
           "log2n": log2n, "n_gpus": world, "steps": steps, "warmup": warmup, "setup_s": round(setup_s, 2),
           "ms_per_proof": round(ms_dev * 1e3, 3), "ms_best": round(best_dev * 1e3, 3),
           "e2e_ms_per_proof": round(ms_e2e * 1e3, 3), "h2d_bytes_per_proof": 32 * n,
           "d2h_bytes_per_proof": 96 * (ell - 1 + 3) + 32 * 3 * ell, "msm_points_per_proof": 4 * n,
           "timing": "wall clock around the host-driven prover (it synchronises at every transcript step), "
                     "max over ranks; transcript = nova_b200.transcript (Keccak-256) on the host",
           "phases_ms": {k: round(v * 1e3, 3) for k, v in (phases or {}).items()}}
    parity = None
    if check_parity:
        parity = {"sharded_equals_single_gpu": None}
        same = proof == proof_e2e
        if world > 1:  # rank 0 proves the whole polynomial on its own GPU: the sharded proof must equal it
            if rank == 0:
                whole = sp.DeviceVec(32 * n)
                check(L.b200_memcpy_h2d(whole.ptr, poly.ctypes.data_as(ctypes.c_void_p), 32 * n))
                single = sp.hyperkzg_prove(curve, ck, whole, x, Keccak256Transcript(p, b"HyperKZG bench"))
                parity["sharded_equals_single_gpu"] = bool(single == proof)
                whole.free()
        if rank == 0:
            parity.update(_check_hyperkzg(ck, poly, x, proof, p, n))
            parity["deterministic"] = bool(same)
            parity["ok"] = bool(parity["verifier_accepts"] and parity["verifier_rejects_tampered"] and same
                                and parity["sharded_equals_single_gpu"] in (None, True))
    _barrier()
    check(L.b200_host_free(pinned))
    ck.release()
    out["parity"] = parity
    out["parity_checked"] = bool(parity and parity.get("ok"))
    return out


def _check_hyperkzg(ck, poly, x, proof, p, n):
    """rank 0, outside every timed region: C and y from the C oracle, then the restated verifier."""
    from oracle import coracle as co
    from oracle import hyperkzg_ref as hk
    from oracle.pyref import CURVES, Keccak256Transcript as OracleTranscript, from_mont_bytes, mont_bytes
    c = CURVES[0]
    t0 = time.time()
    bases = ck.export_bases()
    pb = poly.tobytes()
    C = c.affine_from_bytes(co.msm(0, pb, bases))
    y = from_mont_bytes(p, co.mle_eval(c.scalar_field, pb, b"".join(mont_bytes(p, v) for v in x)))
    oracle_s = time.time() - t0
    com, w, v = proof
    ok = hk.verify(0, TAU, C, x, y, (com, w, v), OracleTranscript(p, b"HyperKZG bench"))
    bad_v = [list(row) for row in v]
    bad_v[len(v) // 2][1] = (bad_v[len(v) // 2][1] + 1) % p
    rej1 = not hk.verify(0, TAU, C, x, y, (com, w, bad_v), OracleTranscript(p, b"HyperKZG bench"))
    rej2 = not hk.verify(0, TAU, C, x, (y + 1) % p, (com, w, v), OracleTranscript(p, b"HyperKZG bench"))
    h = hashlib.sha256(repr((com, w, v)).encode()).hexdigest()[:16]
    return {"verifier_accepts": bool(ok), "verifier_rejects_tampered": bool(rej1 and rej2), "proof_sha256_16": h,
            "checker": "oracle/hyperkzg_ref.verify (hyperkzg.rs:1119-1242, L = [tau]R); C = commit(P) and y = P(x) by "
                       "the C oracle on the exported key", "oracle_seconds": round(oracle_s, 1)}


def hyperkzg_cpu(log2n=20):
    from hyperkzg_replay import cpu
    return cpu(log2n)


# ----------------------------------------------------------------------------------------------------------
# ppsnark (configs[4])
# ----------------------------------------------------------------------------------------------------------
def ppsnark(log2cons=16, steps=2, warmup=1, check_parity=True, device_transcript=True):
    import nova_b200 as nb
    from nova_b200 import fields, ppsnark as dp, spartan as sp
    from nova_b200.native import check, lib
    from nova_b200.transcript import Keccak256Transcript
    from ppsnark_replay import synth_matrix
    L = lib()
    curve = nb.Curve(0)
    fid = curve.scalar_field
    p = fields.MODULUS[fid]
    rng = np.random.default_rng(5)
    m = 1 << log2cons
    num_cons = num_vars = m
    num_io = 2
    ncols = num_vars + 1 + num_io
    t0 = time.time()
    table = np.frombuffer(b"".join(fields.to_mont_bytes(fid, v) for v in (1, p - 1, 2)), dtype=np.uint64).reshape(3, 4)
    mats, rows_all, cols_all, vals = {}, [], [], []
    for name, extra in (("A", 0.6), ("B", 0.3), ("C", 0.1)):
        r, idx, ptr, codes = synth_matrix(rng, num_cons, ncols, extra)
        v = np.ascontiguousarray(table[codes])
        h = ctypes.c_uint64(0)
        check(L.b200_spmv_register(fid, v.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                   ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), num_cons, ncols, ctypes.byref(h)))
        mm = sp.SparseMatrix.__new__(sp.SparseMatrix)
        mm.fid, mm.rows, mm.cols, mm.handle = fid, num_cons, ncols, h.value
        mats[name] = mm
        rows_all.append(r)
        cols_all.append(idx.astype(np.uint32))
        vals.append(v)
    spark = dp.SparkRepr.from_numpy(fid, np.concatenate(rows_all), np.concatenate(cols_all), vals, num_cons, num_vars)
    N = spark.N
    ck = nb.CommitmentKey.setup_tau(curve, N, TAU)
    bits = rng.integers(0, 2, size=num_vars, dtype=np.uint64)  # witness: 90 % bits, 10 % full-width values
    Wd = dp.dev_from_u64(fid, bits)
    wide = rng.integers(0, 1 << 62, size=(num_vars, 4), dtype=np.uint64)
    wide[:, 3] &= np.uint64((1 << 60) - 1)
    sel = np.flatnonzero(rng.random(num_vars) < 0.1)
    if len(sel):
        wb = np.frombuffer(Wd.to_bytes(), dtype=np.uint64).reshape(num_vars, 4).copy()
        wb[sel] = wide[sel]
        check(L.b200_memcpy_h2d(Wd.ptr, wb.ctypes.data_as(ctypes.c_void_p), 32 * num_vars))
    u = int(rng.integers(1, 1 << 62))
    X = [int(rng.integers(1, 1 << 62)) for _ in range(num_io)]
    z = sp.DeviceVec(32 * ncols)  # E = Az o Bz - u Cz: the relaxed instance is satisfied
    check(L.b200_memcpy_d2d(z.ptr, Wd.ptr, 32 * num_vars, None))
    tail = fields.pack(fid, [u] + X)
    check(L.b200_memcpy_h2d(dp.View(z, num_vars).ptr, ctypes.create_string_buffer(tail, len(tail)), len(tail)))
    Az, Bz, Cz = (sp.DeviceVec(32 * num_cons) for _ in range(3))
    for name, o in (("A", Az), ("B", Bz), ("C", Cz)):
        check(L.b200_spmv_dev(mats[name].handle, z.ptr, None, o.ptr, None, None))
    Ed = sp.DeviceVec(32 * num_cons)
    zero, u_dev = dp.dev_zeros(num_cons), dp.dev_scalar(fid, u)
    check(L.b200_cross_term_dev(fid, Az.ptr, Bz.ptr, Cz.ptr, zero.ptr, None, u_dev.ptr, num_cons, Ed.ptr, None))
    check(L.b200_sync())
    U = dict(comm_W=dp.commit_dev(curve, ck, Wd, num_vars), comm_E=dp.commit_dev(curve, ck, Ed, num_cons), u=u, X=X)
    S = dict(num_cons=num_cons, num_vars=num_vars, **mats)
    vk_digest = 4711
    setup_s = time.time() - t0
    runs, proof = [], None
    for rep in range(warmup + steps):
        tm = {}
        check(L.b200_sync())
        t1 = time.perf_counter()
        proof = dp.prove(curve, ck, S, spark, U, dict(W=Wd, E=Ed), vk_digest,
                         Keccak256Transcript(p, b"RelaxedR1CSSNARK"), timings=tm, device_transcript=device_transcript)
        check(L.b200_sync())
        tm["total"] = time.perf_counter() - t1
        if rep >= warmup:
            runs.append(tm)
    mean = sum(t["total"] for t in runs) / len(runs)
    best = min(runs, key=lambda t: t["total"])
    out = {"workload": f"ppsnark RelaxedR1CSSNARK::prove incl. HyperKZG EE::prove, sha256-like synthetic shape, BN254, "
                       f"2^{log2cons} constraints, N = 2^{N.bit_length() - 1}, test SRS [tau^i]G",
           "log2cons": log2cons, "N": N, "n_gpus": 1, "steps": steps, "warmup": warmup, "setup_s": round(setup_s, 2),
           "ms_per_proof": round(mean * 1e3, 3), "ms_best": round(best["total"] * 1e3, 3),
           "device_transcript": bool(device_transcript),
           "phases_ms": {k: round(v * 1e3, 3) for k, v in best.items()},
           "timing": "wall clock around the host-driven prover, device synchronised at the end"}
    parity = None
    if check_parity:
        parity = _check_ppsnark(ck, spark, U, Wd, Ed, num_cons, num_vars, N, vk_digest, proof, p)
    out["parity"] = parity
    out["parity_checked"] = bool(parity and parity.get("ok"))
    ck.release()
    return out


def _check_ppsnark(ck, spark, U, Wd, Ed, num_cons, num_vars, N, vk_digest, proof, p):
    from oracle import coracle as co
    from oracle import ppsnark_ref as pr
    from oracle.pyref import CURVES
    c = CURVES[0]
    t0 = time.time()
    bases = ck.export_bases()

    def commit_ref(dev_vec, n):
        return c.affine_from_bytes(co.msm(0, dev_vec.to_bytes(32 * n), bases[:64 * n]))
    # the instance commitments the prover absorbed must be the oracle's (commit at this size and distribution)
    inst_ok = U["comm_W"] == commit_ref(Wd, num_vars) and U["comm_E"] == commit_ref(Ed, num_cons)
    S_comm = {k: commit_ref(getattr(spark, k), N) for k in ("val_A", "val_B", "val_C", "row", "col", "ts_row", "ts_col")}
    oracle_s = time.time() - t0
    pf = {k: v for k, v in proof.items() if k not in ("batched_poly",)}
    try:
        ok = bool(pr.verify(p, c, 0, TAU, num_cons, num_vars, N, U, S_comm, vk_digest, pf))
        why = ""
    except AssertionError as e:
        ok, why = False, str(e)
    bad = dict(pf, eval_W=(pf["eval_W"] + 1) % p)
    try:
        rej = not pr.verify(p, c, 0, TAU, num_cons, num_vars, N, U, S_comm, vk_digest, bad)
    except AssertionError:
        rej = True
    digest = hashlib.sha256(repr(sorted((k, v) for k, v in pf.items() if k != "transcript")).encode()).hexdigest()[:16]
    return {"ok": bool(ok and rej and inst_ok), "verifier_accepts": ok, "verifier_rejects_tampered": bool(rej),
            "instance_commitments_equal_oracle": bool(inst_ok), "why": why, "proof_sha256_16": digest,
            "checker": "oracle/ppsnark_ref.verify (ppsnark.rs:1386-1655 restated; KZG as L = [tau]R); the 9 instance / "
                       "shape commitments by the C oracle on the exported key", "oracle_seconds": round(oracle_s, 1)}


# ----------------------------------------------------------------------------------------------------------
# prove_step (configs[2])
# ----------------------------------------------------------------------------------------------------------
def prove_step(steps=5, warmup=2, check_parity=True, cpu_steps=1):
    import prove_step_replay as psr
    g = psr.gpu_replay(steps=steps, warmup=warmup, return_outputs=check_parity)
    out = {"workload": "RecursiveSNARK::prove_step kernel-sequence replay, MinRoot-sized shapes, BN254 / Grumpkin",
           "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": round(g["ms_per_step"], 4),
           "primary_constraints": g["primary_constraints"], "secondary_constraints": g["secondary_constraints"],
           "ops_per_step": g["ops_per_step"], "excluded": g["excluded"],
           "h2d_bytes_per_step": g.get("h2d_bytes_per_step"), "d2h_bytes_per_step": g.get("d2h_bytes_per_step")}
    parity = None
    if check_parity:
        c = psr.cpu_replay(steps=cpu_steps, return_outputs=True)
        same = g["outputs"] == c["outputs"]
        parity = {"ok": bool(same), "gpu": g["outputs"], "cpu": c["outputs"] if not same else "identical",
                  "checker": "the same op sequence through oracle/oracle.c: 4 commitments (affine) and SHA-256 of the "
                             "4 folded vectors must be equal"}
        out["cpu_baseline"] = {k: v for k, v in c.items() if k != "outputs"}
    out["parity"] = parity
    out["parity_checked"] = bool(parity and parity["ok"])
    return out


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["hyperkzg", "ppsnark", "prove_step"])
    ap.add_argument("--log2", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    from nova_b200.native import check, lib
    check(lib().b200_init(0))
    if a.workload == "hyperkzg":
        print(json.dumps(hyperkzg(a.log2 or 20, steps=a.steps)))
    elif a.workload == "ppsnark":
        print(json.dumps(ppsnark(a.log2 or 14, steps=a.steps)))
    else:
        print(json.dumps(prove_step(steps=a.steps)))
