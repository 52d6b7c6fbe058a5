#!/usr/bin/env python3
"""Timing probe: the device Poseidon RO squeeze (9 elements, wide sponge: the NIFS challenge) against the host round trip
it replaces (D2H of 96 bytes + H2D of 32 bytes around a host hash)."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from nova_b200 import fields, poseidon as dp
from nova_b200.native import check, lib
from nova_b200.spartan import DeviceVec

L = lib()
check(L.b200_init(0))
for fid, arity, n in ((1, 24, 9), (0, 24, 9), (1, 5, 9)):
    c = dp.PoseidonConstants.get(fid, arity)
    d = DeviceVec.from_bytes(fields.pack(fid, list(range(1, n + 1))))
    out = DeviceVec(96)
    for _ in range(3):
        check(L.b200_poseidon_ro_dev(c.handle, d.ptr, n, 128, 0, out.ptr, None))
    check(L.b200_sync())
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        check(L.b200_poseidon_ro_dev(c.handle, d.ptr, n, 128, 0, out.ptr, None))
    check(L.b200_sync())
    dev_us = (time.perf_counter() - t0) / reps * 1e6
    host = ctypes.create_string_buffer(96)
    t0 = time.perf_counter()
    for _ in range(reps):
        check(L.b200_memcpy_d2h(host, out.ptr, 96))
        check(L.b200_memcpy_h2d(out.ptr, host, 32))
    rt_us = (time.perf_counter() - t0) / reps * 1e6
    print({"field": fid, "arity": arity, "elements": n, "rounds": (c.r_f, c.r_p), "device_squeeze_us": round(dev_us, 1),
           "host_round_trip_us_without_the_hash": round(rt_us, 1)})
