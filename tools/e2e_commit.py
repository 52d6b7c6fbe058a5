#!/usr/bin/env python3
"""End-to-end commit timing without torch: b200_commit of 2^k BN254 scalars from PINNED host memory
(b200_host_alloc), wall clock around blocking calls -- the same quantity bench.py reports as `e2e`, for quick A/B
of host-path tuning hooks (they are read once per process, so run the script once per setting):

    python tools/e2e_commit.py --log-n 20
    NOVA_B200_H2D_CHUNKS=4 python tools/e2e_commit.py --log-n 20
"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import nova_b200 as nb  # noqa: E402
from nova_b200.native import check, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    L = lib()
    check(L.b200_init(0))
    n = 1 << a.log_n
    ck = nb.CommitmentKey.setup_synthetic(nb.Curve(0), n)
    host = ctypes.c_void_p()
    check(L.b200_host_alloc(32 * n, ctypes.byref(host)))
    import random
    rnd = random.Random(1)
    block = b"".join(rnd.getrandbits(250).to_bytes(32, "little") for _ in range(1 << 12))
    for off in range(0, 32 * n, len(block)):
        ctypes.memmove(host.value + off, block, min(len(block), 32 * n - off))
    out = ctypes.create_string_buffer(96)
    for _ in range(3):
        check(L.b200_commit(ck.handle, host, n, None, out))
    first = out.raw
    t0 = time.perf_counter()
    for _ in range(a.reps):
        check(L.b200_commit(ck.handle, host, n, None, out))
    dt = (time.perf_counter() - t0) / a.reps
    print(json.dumps({"what": "b200_commit from pinned host memory", "log_n": a.log_n, "ms": round(dt * 1e3, 4),
                      "pairs_per_s": round(n / dt), "h2d_chunks": os.environ.get("NOVA_B200_H2D_CHUNKS", "1"),
                      "stable": out.raw == first, "result_head": first[:8].hex()}))
    check(L.b200_host_free(host))


if __name__ == "__main__":
    main()
