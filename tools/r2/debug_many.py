#!/usr/bin/env python3
"""Debug probe: b200_commit_many_dev (lanes) against host commits, small keys (pool on/off via NOVA_B200_POOL)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import nova_b200 as nb
from nova_b200.native import check, lib
from nova_b200.spartan import DeviceVec, commit_many_dev
from oracle import coracle as co
from oracle.pyref import CURVES

check(lib().b200_init(0))
cid, c = 0, CURVES[0]
print("pool env:", os.environ.get("NOVA_B200_POOL"))
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << logn
ck = nb.CommitmentKey.setup_synthetic(nb.Curve(cid), n)
ce = nb.CommitmentEngine(cid)
sc = co.gen_scalars(c.scalar_field, 4242, n)
for rep in range(3):
    lens = [n, n // 2, 700, 4096 % (n + 1), 33, 2, 1, 1000 % (n + 1), 5, n - 3]
    vecs = [DeviceVec.from_bytes(sc[32 * 3:32 * (3 + m)] if 3 + m <= n else sc[:32 * m]) for m in lens]
    many = commit_many_dev(cid, ck, vecs, lens)
    bad = []
    for m, got in zip(lens, many):
        exp = ce.commit(ck, sc[32 * 3:32 * (3 + m)] if 3 + m <= n else sc[:32 * m], None)
        if got != exp:
            bad.append(m)
    print(f"rep {rep} key 2^{logn}: mismatching lengths:", bad)
