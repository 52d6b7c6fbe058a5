#!/usr/bin/env python3
"""The table the reference's `commit` bench would print (benches/commit.rs:32-110, 112-249): BN254
commitments of n scalars drawn as u1 / u10 / u16 / u32 / u64 / uniform field elements, through
 (a) CE::commit on field scalars      -> b200_commit   (host pointers, pinned, H2D inside)
 (b) CE::commit_small on integers     -> b200_msm_small (u8/u16/u32/u64 elements, H2D inside)
and, with --cpu, the CPU restatement of msm() / msm_small() on the host cores.
One JSON line per (n, distribution).  Usage: python tools/commit_table.py [--cpu] [log2n ...]"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import nova_b200 as nb
from nova_b200.native import check, lib

R_MONT = (1 << 256) % 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def small_to_mont(vals: np.ndarray) -> np.ndarray:
    """u64 integers -> Montgomery field elements (v * R mod p) as (n,4) u64, vectorised via Python ints
    only for the distinct small values when few, else per element."""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    uniq, inv = np.unique(vals, return_inverse=True)
    if len(uniq) <= 1 << 16:
        tab = np.zeros((len(uniq), 4), dtype=np.uint64)
        for i, v in enumerate(uniq):
            m = int(v) * R_MONT % P
            tab[i] = [(m >> (64 * k)) & ((1 << 64) - 1) for k in range(4)]
        return tab[inv]
    for i, v in enumerate(vals):
        m = int(v) * R_MONT % P
        out[i] = [(m >> (64 * k)) & ((1 << 64) - 1) for k in range(4)]
    return out


def main():
    use_cpu = "--cpu" in sys.argv
    logs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [20]
    L = lib()
    check(L.b200_init(0))
    rng = np.random.default_rng(7)
    for lg in logs:
        n = 1 << lg
        ck = nb.CommitmentKey.setup_synthetic(nb.Curve(0), n)
        host = ctypes.c_void_p()
        check(L.b200_host_alloc(n * 32, ctypes.byref(host)))
        out = ctypes.create_string_buffer(96)
        if use_cpu:
            from oracle import coracle as co
            sys.path.insert(0, ROOT)
            import bench
            cores = bench.effective_cores()
            bases = co.gen_bases(0, n)
        for name, bits, eb in (("u1", 1, 1), ("u10", 10, 2), ("u16", 16, 2), ("u32", 32, 4), ("u64", 64, 8),
                               ("uniform", 254, 0)):
            if bits <= 64:
                hi = (1 << bits) - 1
                vals = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True)
                if lg > 20 and bits > 16:  # Montgomery conversion in Python is slow: sample + tile
                    base = small_to_mont(vals[: 1 << 16])
                    fe = np.tile(base, (n >> 16, 1))
                    vals = np.tile(vals[: 1 << 16], n >> 16)
                else:
                    fe = small_to_mont(vals)
            else:
                fe = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
                fe[:, 3] &= np.uint64((1 << 60) - 1)
            ctypes.memmove(host, fe.ctypes.data, n * 32)
            rec = {"log2n": lg, "scalars": name}
            for _ in range(2):
                check(L.b200_commit(ck.handle, host, n, None, out))
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                check(L.b200_commit(ck.handle, host, n, None, out))
            rec["commit_ms"] = round((time.perf_counter() - t0) * 1e3 / reps, 3)
            ref_commit = out.raw
            if eb:
                small = vals.astype({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[eb])
                sbuf = small.ctypes.data_as(ctypes.c_void_p)
                for _ in range(2):
                    check(L.b200_msm_small(ck.handle, 0, sbuf, eb, n, bits, out))
                t0 = time.perf_counter()
                for _ in range(reps):
                    check(L.b200_msm_small(ck.handle, 0, sbuf, eb, n, bits, out))
                rec["commit_small_ms"] = round((time.perf_counter() - t0) * 1e3 / reps, 3)
                # both entry points must give the same group element (compare affine via cross-multiplying)
                rec["small_equals_field"] = _same_point(ref_commit, out.raw)
            if use_cpu:
                sc = fe.tobytes()
                co.msm(0, sc, bases, cores)
                t0 = time.perf_counter()
                co.msm(0, sc, bases, cores)
                rec["cpu_commit_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
                rec["cpu_cores"] = cores
            print(json.dumps(rec), flush=True)
        check(L.b200_host_free(host))
        ck.release()


def _same_point(j1: bytes, j2: bytes) -> bool:
    Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
    def coords(j):
        return [int.from_bytes(j[i:i + 32], "little") for i in (0, 32, 64)]
    x1, y1, z1 = coords(j1)
    x2, y2, z2 = coords(j2)
    if z1 == 0 or z2 == 0:
        return z1 == z2
    # Montgomery factors cancel in the cross products up to a common power of R
    return (x1 * z2 * z2 - x2 * z1 * z1) % Q == 0 and (y1 * z2 ** 3 - y2 * z1 ** 3) % Q == 0


if __name__ == "__main__":
    main()
