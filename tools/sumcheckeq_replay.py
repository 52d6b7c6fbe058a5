#!/usr/bin/env python3
"""Replay of the reference's `sumcheckeq` bench (benches/sumcheckeq.rs:24-115): MemorySumcheckInstance::new
over deterministic vectors v_k[i] = i * k (k = 1..8; ts_row = v_1, ts_col = v_2), taus[i] = -2 i, then per round
`evaluation_points()` followed by `bound(r_i)` with r_i = -i.  (taus[0] = 0, so round 1 always takes the
third-sum fall-back.)  Times the whole loop on device-resident vectors for lengths 2^3 .. 2^max.

    python tools/sumcheckeq_replay.py [--min 3] [--max 22] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def run_sc(num_vars: int, fid: int = 0, collect: bool = False):
    """One benchmark iteration; returns (seconds, evaluation points per round if `collect`)."""
    import numpy as np
    from nova_b200 import fields, ppsnark as dp
    from nova_b200.native import check, lib
    from nova_b200.spartan import _challenge_dev
    p = fields.MODULUS[fid]
    n = 1 << num_vars
    idx = np.arange(n, dtype=np.uint64)
    vs = [dp.dev_from_u64(fid, idx * np.uint64(k)) for k in range(1, 9)]
    rs = [(-i) % p for i in range(num_vars)]
    taus = [(-2 * i) % p for i in range(num_vars)]
    check(lib().b200_sync())
    t0 = time.perf_counter()
    inst = dp.MemorySumcheckInstance(fid, n, vs[0:4], vs[4:8], taus, vs[0], vs[1])
    sums = dp.RoundSums(fid)
    out = []
    for r in rs:
        inst.enqueue(sums)
        ev = inst.evaluation_points(sums.fetch())
        if collect:
            out.append(ev)
        inst.bound(r, _challenge_dev(fid, r))
    check(lib().b200_sync())
    return time.perf_counter() - t0, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=3)
    ap.add_argument("--max", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from nova_b200.native import check, lib
    check(lib().b200_init(0))
    for nv in range(a.min, a.max + 1):
        run_sc(nv)  # warm-up
        best = min(run_sc(nv)[0] for _ in range(a.reps))
        print(json.dumps({"bench": f"NovaProve-PPSNARK-SumCheckEq-len-{nv}/ProveMemory", "ms": round(best * 1e3, 4),
                          "rounds": nv, "includes": "MemorySumcheckInstance::new (copies) + rounds"}), flush=True)


if __name__ == "__main__":
    main()
