#!/usr/bin/env python3
"""Debug probe (round 2): host-pointer commit (sliced or single-shot, NOVA_B200_E2E_SLICES) and b200_jacobian_sum_dev
against the closed form / the oracle's group law."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import nova_b200 as nb
from nova_b200.native import check, lib
from nova_b200.provider import Curve, _cbuf, _jac_to_affine
from nova_b200.spartan import DeviceVec
from oracle import coracle as co
from oracle.pyref import CURVES

L = lib()
check(L.b200_init(0))
cid, c = 0, CURVES[0]
K0 = 0x5EED
print("slices env:", os.environ.get("NOVA_B200_E2E_SLICES"))
# 1. jacobian sum of k points
g = c.affine_bytes(c.gen)
for k in (1, 2, 3, 4, 5, 8, 9, 17):
    pts, exp = b"", None
    for i in range(k):
        a = co.scalar_mul(cid, g, 1000 + 7 * i)
        P = c.affine_from_bytes(a)
        exp = c.add(exp, P)
        pts += a + (1 << 256).to_bytes(33, "little")[:0] + ((1 << 256) % c.p).to_bytes(32, "little")  # z = 1 (Montgomery)
    d = DeviceVec.from_bytes(pts)
    out = DeviceVec(96)
    check(L.b200_jacobian_sum_dev(cid, d.ptr, k, out.ptr, None))
    got = _jac_to_affine(Curve(cid), out.to_bytes())
    print("jacobian_sum k =", k, "ok" if got == exp else "MISMATCH")
# 2. host commits vs the closed form
for keylog, sizes in ((20, [(1 << 19) + 5, 1 << 20]), (22, [(1 << 19) + 5, (1 << 21) + 7, 1 << 22])):
    ck = nb.CommitmentKey.setup_synthetic(nb.Curve(cid), 1 << keylog, k0=K0)
    for n in sizes:
        sc = co.gen_scalars(c.scalar_field, 11, n)
        out = ctypes.create_string_buffer(96)
        check(L.b200_commit(ck.handle, _cbuf(sc), n, None, out))
        got = _jac_to_affine(Curve(cid), out.raw)
        k = co.dot_index(c.scalar_field, sc, K0)
        exp = c.affine_from_bytes(co.scalar_mul(cid, g, k))
        d = DeviceVec.from_bytes(sc)
        o2 = DeviceVec(96)
        check(L.b200_msm_dev(ck.handle, 0, d.ptr, n, o2.ptr, None))
        got2 = _jac_to_affine(Curve(cid), o2.to_bytes())
        print(f"key 2^{keylog} n = {n}: host commit", "ok" if got == exp else "MISMATCH", "| msm_dev", "ok" if got2 == exp else "MISMATCH")
    ck.release()
