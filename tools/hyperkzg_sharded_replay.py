#!/usr/bin/env python3
"""HyperKZG EvaluationEngine::prove over N GPUs (SURVEY.md §8d config C4 "sharded 1/2/4/8 GPUs"):
nova_b200.sharding.sharded_hyperkzg_prove on a uniformly random polynomial of 2^LOG2N BN254 scalars split by index
range, one process per GPU.  Wall time per proof, max over ranks, after one warm-up proof; the digest printed is
the one tools/hyperkzg_replay.py prints for the same size (same key, polynomial and challenges), so the two
tools check each other.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \\
        tools/hyperkzg_sharded_replay.py --log2n 22 [--comm nccl|host] [--reps 3]

--comm nccl (default): device-to-device all-gathers over NVLink (NcclComm); host: staged through the host
(HostStagedComm; any backend).  Without torchrun it runs as a single rank.
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--comm", choices=["nccl", "host"], default="nccl")
    a = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl" if a.comm == "nccl" else "gloo", rank=rank, world_size=world)
    import nova_b200 as nb
    from hyperkzg_replay import K0, challenges, synth_poly
    from nova_b200 import sharding as sh
    from nova_b200 import spartan as sp
    from nova_b200.native import check, lib
    L = lib()
    check(L.b200_init(local))
    n = 1 << a.log2n
    curve = nb.Curve(0)
    ck = nb.CommitmentKey.setup_synthetic(curve, n, k0=K0)  # the full key on every GPU
    lo, hi = rank * (n // world), (rank + 1) * (n // world)
    poly = synth_poly(n, 4)[lo:hi].copy()
    P = sp.DeviceVec(32 * (hi - lo))
    check(L.b200_memcpy_h2d(P.ptr, poly.ctypes.data_as(ctypes.c_void_p), 32 * (hi - lo)))
    x, r, q = challenges(a.log2n)
    comm = sh.NcclComm() if (world > 1 and a.comm == "nccl") else sh.HostStagedComm()
    times = []
    for rep in range(a.reps + 1):
        if world > 1:
            dist.barrier()
        check(L.b200_sync())
        t0 = time.perf_counter()
        com, v, w = sh.sharded_hyperkzg_prove(curve, ck, P, x, r, q, comm)
        check(L.b200_sync())
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda" if a.comm == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if rep:
            times.append(dt)
    out = None
    if rank == 0:
        out = ({"workload": f"HyperKZG prove core, BN254, 2^{a.log2n} uniform scalars, {world} GPU(s), "
                                      f"index-range sharding, comm={a.comm if world > 1 else 'none'}",
                          "n_gpus": world, "log2n": a.log2n, "ms_best": round(min(times) * 1e3, 3),
                          "ms_all": [round(t * 1e3, 3) for t in times], "timing": "wall clock, max over ranks",
                          "digest": [com[0][0] % (1 << 64), w[2][0] % (1 << 64)]})
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
