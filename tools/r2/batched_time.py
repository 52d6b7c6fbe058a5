"""Times the batched sum-check (prove_helper, ppsnark.rs:886-983) on random tables: the one-call loop
(b200_sumcheck_batched) per tail setting against the per-launch loop (prove_helper_device_rounds)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nova_b200 import fields, ppsnark as dp, spartan as sp  # noqa: E402
from nova_b200.native import check, lib  # noqa: E402
from oracle.pyref import FIELD_MODULUS, Keccak256Transcript, SplitMix64  # noqa: E402


def main():
    fid = 0
    p = FIELD_MODULUS[fid]
    check(lib().b200_init(0))
    for ell in [int(a) for a in sys.argv[1:]] or [14, 18, 20]:
        N = 1 << ell
        rng = np.random.default_rng(ell)

        def vec():
            v = rng.integers(0, 1 << 62, size=(N, 4), dtype=np.uint64)
            v[:, 3] &= np.uint64((1 << 59) - 1)
            return sp.DeviceVec.from_bytes(v.tobytes())
        r = SplitMix64(ell)
        rhos, r_outer = [r.field(p) for _ in range(ell)], [r.field(p) for _ in range(ell)]
        claim, claim_E = r.field(p), r.field(p)
        base = [vec() for _ in range(15)]

        def run(helper):
            cp = [dp.dev_copy(v, N) for v in base]
            mem = dp.MemorySumcheckInstance(fid, N, cp[0:4], cp[4:8], rhos, cp[8], cp[9])
            inner = dp.InnerBatchedSumcheckInstance(fid, N, claim, cp[10], cp[11], cp[12], claim_E, r_outer, cp[13])
            wit = dp.WitnessBoundSumcheck(fid, N, r_outer, cp[14], 1 << (ell - 1))
            tr = Keccak256Transcript(p, b"zt")
            tr.absorb_scalar(b"k", 3)
            check(lib().b200_sync())
            t0 = time.perf_counter()
            out = helper(fid, mem, inner, wit, tr)
            check(lib().b200_sync())
            return (time.perf_counter() - t0) * 1e3, out[0][-1], tr.squeeze(b"x")
        ref = None
        for name, helper, tb in [("rounds", dp.prove_helper_device_rounds, None)] + [("one-call tail=%d" % t, dp.prove_helper_device, t) for t in (0, 8)]:
            if tb is not None:
                lib().b200_sumcheck_tail_bits(tb)
            ts = []
            for _ in range(4):
                ms, last, sq = run(helper)
                ts.append(ms)
                ref = ref or (last, sq)
                assert (last, sq) == ref, name
            print(f"ell={ell} {name:20s} best {min(ts):8.3f} ms  ({min(ts) * 1e3 / ell:6.1f} us/round)", flush=True)
        lib().b200_sumcheck_tail_bits(8)


main()
