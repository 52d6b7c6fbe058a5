#!/usr/bin/env python3
"""HBM-roofline probe for the streaming field kernels (device-resident, CUDA events on the
launching stream).  Prints one JSON line per kernel: algorithmic bytes, ms, GB/s, fraction of
the measured HBM peak (MEASURED_PEAKS.json).  Usage: python tools/fieldbench.py [log2n]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from nova_b200.native import check, lib


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << lg
    L = lib()
    check(L.b200_init(0))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    fid = 0
    rng = np.random.default_rng(1)

    def vec(m):
        a = rng.integers(0, 1 << 62, size=(m, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()

    A, B, C, E, T = vec(n), vec(n), vec(n), vec(n), vec(n)
    u = vec(1)
    out3 = torch.zeros(96, dtype=torch.uint8, device="cuda")
    eqr = vec(n // 2)
    stream = torch.cuda.Stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    cases = [
        ("cross_term (T = Az*Bz - u*Cz - E)", 160 * n,
         lambda: L.b200_cross_term_dev(fid, P(A), P(B), P(C), P(E), None, P(u), n, P(T), sp)),
        ("axpy (fold W1 + r*W2)", 96 * n, lambda: L.b200_axpy_dev(fid, P(A), P(B), P(u), n, P(T), sp)),
        ("vec_add (Z1 + Z2)", 96 * n, lambda: L.b200_vec_add_dev(fid, P(A), P(B), n, P(T), sp)),
        ("bind_poly_var_top", 96 * (n // 2), lambda: L.b200_bind_top_dev(fid, P(T), n, P(u), sp)),
        ("sumcheck eq_cubic3 round (t0,tinf)", 160 * (n // 2) + 32 * (n // 2),
         lambda: L.b200_sc_eval_dev(fid, 4, P(A), P(B), P(C), n, None, P(eqr), 0, P(out3), sp)),
        ("sumcheck quad_prod round", 128 * (n // 2),
         lambda: L.b200_sc_eval_dev(fid, 0, P(A), P(B), None, n, None, None, 0, P(out3), sp)),
        ("kzg_fold", 96 * (n // 2), lambda: L.b200_kzg_fold_dev(fid, P(A), n, P(u), P(T), sp)),
        ("poly_eval x3 (Horner)", 32 * n, lambda: L.b200_poly_eval_dev(fid, P(A), n, P(eqr), 3, P(out3), sp)),
        ("poly_div (X - u)", 96 * n, lambda: L.b200_poly_div_dev(fid, P(A), n, P(u), P(T), sp)),
    ]
    with torch.cuda.stream(stream):
        for name, nbytes, fn in cases:
            for _ in range(3):
                check(fn())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record(stream)
            for _ in range(iters):
                check(fn())
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"kernel": name, "log2n": lg, "algorithmic_bytes": nbytes, "ms": round(ms, 4),
                              "achieved_GBps": round(gbs, 1), "peak_GBps": peak, "frac": round(gbs / peak, 3)}),
                  flush=True)


if __name__ == "__main__":
    main()
