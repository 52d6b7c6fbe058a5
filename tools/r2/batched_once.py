"""One batched sum-check (b200_sumcheck_batched) on random tables, for a launch list: python batched_once.py ELL TAIL_BITS"""
import sys

import numpy as np

sys.path.insert(0, ".")
from nova_b200 import ppsnark as dp, spartan as sp  # noqa: E402
from nova_b200.native import check, lib  # noqa: E402
from oracle.pyref import FIELD_MODULUS, Keccak256Transcript, SplitMix64  # noqa: E402

ell, tail = int(sys.argv[1]), int(sys.argv[2])
fid = 0
p = FIELD_MODULUS[fid]
check(lib().b200_init(0))
lib().b200_sumcheck_tail_bits(tail)
N = 1 << ell
rng = np.random.default_rng(ell)


def vec():
    v = rng.integers(0, 1 << 62, size=(N, 4), dtype=np.uint64)
    v[:, 3] &= np.uint64((1 << 59) - 1)
    return sp.DeviceVec.from_bytes(v.tobytes())


r = SplitMix64(ell)
rhos, r_outer = [r.field(p) for _ in range(ell)], [r.field(p) for _ in range(ell)]
for rep in range(2):
    cp = [vec() for _ in range(15)]
    mem = dp.MemorySumcheckInstance(fid, N, cp[0:4], cp[4:8], rhos, cp[8], cp[9])
    inner = dp.InnerBatchedSumcheckInstance(fid, N, 5, cp[10], cp[11], cp[12], 7, r_outer, cp[13])
    wit = dp.WitnessBoundSumcheck(fid, N, r_outer, cp[14], 1 << (ell - 1))
    tr = Keccak256Transcript(p, b"zt")
    check(lib().b200_sync())
    print("MARK begin", rep, flush=True)
    dp.prove_helper_device(fid, mem, inner, wit, tr)
    check(lib().b200_sync())
print("done")
