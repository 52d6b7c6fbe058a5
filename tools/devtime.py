#!/usr/bin/env python3
"""Developer timing probe (not the benchmark): per-stage and whole-MSM device times at a few sizes
and window widths, with CUDA events on torch's current stream.  Usage on the GPU box:
    python tools/devtime.py [log2n ...]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import nova_b200 as nb
from nova_b200.native import check, lib
from oracle import coracle as co
from oracle.pyref import CURVES


def time_msm(ck, d_scalars, n, d_out, iters=5):
    L = lib()
    s = torch.cuda.Stream()  # a real (non-NULL) stream: NULL would select the library's own
    sp = ctypes.c_void_p(s.cuda_stream)
    torch.cuda.synchronize()
    for _ in range(2):
        check(L.b200_msm_dev(ck.handle, 0, d_scalars.data_ptr(), n, d_out.data_ptr(), sp))
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record(s)
    for i in range(iters):
        check(L.b200_msm_dev(ck.handle, 0, d_scalars.data_ptr(), n, d_out.data_ptr(), sp))
        ev[i + 1].record(s)
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return min(ts), sum(ts) / len(ts)


def main():
    check(lib().b200_init(0))
    logs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [16, 20]
    wins = [int(a[2:]) for a in sys.argv[1:] if a.startswith("c=")] or [0]
    dist = ([a[5:] for a in sys.argv[1:] if a.startswith("dist=")] or ["uniform"])[0]
    cid = 0
    c = CURVES[cid]
    for lg in logs:
        n = 1 << lg
        t0 = time.time()
        bases = co.gen_bases(cid, n)
        sc = co.gen_scalars(c.scalar_field, 2, n)
        if dist == "half_equal":  # ppsnark-style padding: the upper half repeats one full-width value
            sc = sc[:16 * n] + sc[:32] * (n - n // 2)
        elif dist == "bits":      # 0/1 witness
            import random
            rnd = random.Random(1)
            one = co.field_from_u64(c.scalar_field, [1])
            sc = b"".join(one if rnd.random() < 0.5 else bytes(32) for _ in range(n))
        print(f"  scalar distribution: {dist}")
        d_sc = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
        d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
        print(f"n=2^{lg}: inputs generated in {time.time()-t0:.1f}s", flush=True)
        for w in wins:
            t0 = time.time()
            ck = nb.CommitmentKey(nb.Curve(cid), bases, None, w)
            treg = time.time() - t0
            best, avg = time_msm(ck, d_sc, n, d_out)
            cc = ctypes.c_int(0); nt = ctypes.c_int(0)
            lib().b200_ck_len(ck.handle, None, ctypes.byref(cc), ctypes.byref(nt))
            L = lib()
            check(L.b200_profile_enable(1)); check(L.b200_profile_reset())
            time_msm(ck, d_sc, n, d_out, iters=3)
            st = (ctypes.c_double * 5)(); mm = ctypes.c_uint64(0); ll = ctypes.c_uint64(0)
            check(L.b200_profile_read(st, 5, ctypes.byref(mm), ctypes.byref(ll)))
            check(L.b200_profile_enable(0))
            stages = " ".join(f"{nm}={st[i] / max(1, mm.value):.3f}" for i, nm in
                              enumerate(["digits", "sort", "acc", "fixup", "reduce"]))
            print(f"  stages(ms): {stages}")
            print(f"  c={cc.value} tables={nt.value} register {treg*1e3:.0f} ms | msm best {best:.3f} ms avg {avg:.3f} ms "
                  f"-> {n/best/1e3:.1f} M pairs/s", flush=True)
            ck.release()


if __name__ == "__main__":
    main()
