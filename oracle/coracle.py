"""ctypes binding of oracle/liboracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _buf(b: bytes):
    return ctypes.create_string_buffer(bytes(b), len(b))


def ncores() -> int:
    return os.cpu_count() or 1


def fe_op(fid: int, op: int, a: bytes, b: bytes | None = None) -> bytes:
    n = len(a) // 32
    out = ctypes.create_string_buffer(32 * n)
    rc = lib().orc_fe_op(fid, op, _buf(a), _buf(b if b is not None else a), out, ctypes.c_size_t(n))
    assert rc == 0
    return out.raw


def jacobian_to_affine(curve: int, jac: bytes) -> bytes:
    n = len(jac) // 96
    out = ctypes.create_string_buffer(64 * n)
    assert lib().orc_jacobian_to_affine(curve, _buf(jac), ctypes.c_size_t(n), out) == 0
    return out.raw


def _msm_call(fn, curve, scalars, bases, n, nthreads):
    out = ctypes.create_string_buffer(64)
    rc = fn(curve, scalars, bases, ctypes.c_size_t(n), nthreads, out)
    assert rc == 0
    return out.raw


def msm(curve: int, scalars: bytes, bases: bytes, nthreads: int = 0) -> bytes:
    """src/provider/msm.rs:225 `msm` restatement -> affine 64 B (zeros = identity)."""
    n = len(scalars) // 32
    assert len(bases) >= 64 * n
    return _msm_call(lib().orc_msm, curve, _buf(scalars), _buf(bases), n, nthreads or ncores())


def msm_best(curve: int, scalars: bytes, bases: bytes, nthreads: int = 0) -> bytes:
    n = len(scalars) // 32
    return _msm_call(lib().orc_msm_best, curve, _buf(scalars), _buf(bases), n, nthreads or ncores())


def msm_naive(curve: int, scalars: bytes, bases: bytes, nthreads: int = 0) -> bytes:
    n = len(scalars) // 32
    return _msm_call(lib().orc_msm_naive, curve, _buf(scalars), _buf(bases), n, nthreads or ncores())


def msm_small(curve: int, scalars_u64, bases: bytes, max_bits: int = -1, nthreads: int = 0) -> bytes:
    n = len(scalars_u64)
    arr = (ctypes.c_uint64 * max(n, 1))(*scalars_u64)
    out = ctypes.create_string_buffer(64)
    rc = lib().orc_msm_small(curve, arr, _buf(bases), ctypes.c_size_t(n), max_bits,
                             nthreads or ncores(), out)
    assert rc == 0
    return out.raw


def batch_add(curve: int, bases: bytes, idx, nthreads: int = 0) -> bytes:
    m = len(idx)
    arr = (ctypes.c_uint64 * max(m, 1))(*idx)
    out = ctypes.create_string_buffer(64)
    assert lib().orc_batch_add(curve, _buf(bases), arr, ctypes.c_size_t(m), nthreads or ncores(), out) == 0
    return out.raw


def cross_term(fid, az, bz, cz, e1, e2, u) -> bytes:
    n = len(az) // 32
    out = ctypes.create_string_buffer(32 * n)
    rc = lib().orc_cross_term(fid, _buf(az), _buf(bz), _buf(cz), _buf(e1),
                              _buf(e2) if e2 is not None else None, _buf(u), ctypes.c_size_t(n), out)
    assert rc == 0
    return out.raw


def axpy(fid, a, b, r) -> bytes:
    n = len(a) // 32
    out = ctypes.create_string_buffer(32 * n)
    assert lib().orc_axpy(fid, _buf(a), _buf(b), _buf(r), ctypes.c_size_t(n), out) == 0
    return out.raw


def vec_add(fid, a, b) -> bytes:
    n = len(a) // 32
    out = ctypes.create_string_buffer(32 * n)
    assert lib().orc_vec_add(fid, _buf(a), _buf(b), ctypes.c_size_t(n), out) == 0
    return out.raw


def bind_top(fid, z, r) -> bytes:
    n = len(z) // 32
    zb = _buf(z)
    assert lib().orc_bind_top(fid, zb, ctypes.c_size_t(n), _buf(r)) == 0
    return zb.raw[: 32 * (n // 2)]


def field_from_u64(fid, vals) -> bytes:
    n = len(vals)
    arr = (ctypes.c_uint64 * max(n, 1))(*vals)
    out = ctypes.create_string_buffer(32 * n)
    assert lib().orc_field_from_u64(fid, arr, ctypes.c_size_t(n), out) == 0
    return out.raw


# ---- synthetic inputs (shared by tests and bench; SplitMix64 streams match pyref) -------------
K0_DEFAULT = 0x5EED


def gen_scalars(fid: int, seed: int, n: int) -> bytes:
    out = ctypes.create_string_buffer(max(32 * n, 1))
    assert lib().orc_gen_scalars(fid, ctypes.c_uint64(seed), ctypes.c_size_t(n), out) == 0
    return out.raw[: 32 * n]


def _limbs(k: int):
    return (ctypes.c_uint64 * 4)(*[(k >> (64 * i)) & ((1 << 64) - 1) for i in range(4)])


def gen_bases(curve: int, n: int, k0: int = K0_DEFAULT) -> bytes:
    from . import pyref
    c = pyref.CURVES[curve]
    out = ctypes.create_string_buffer(max(64 * n, 1))
    assert lib().orc_gen_bases(curve, _buf(c.affine_bytes(c.gen)), _limbs(k0), ctypes.c_size_t(n), out) == 0
    return out.raw[: 64 * n]


def dot_index(fid: int, scalars: bytes, k0: int = K0_DEFAULT) -> int:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(32)
    assert lib().orc_dot_index(fid, _buf(scalars), ctypes.c_size_t(n), _limbs(k0), out) == 0
    return int.from_bytes(out.raw, "little")


def scalar_mul(curve: int, pt: bytes, k: int) -> bytes:
    out = ctypes.create_string_buffer(64)
    assert lib().orc_scalar_mul(curve, _buf(pt), _limbs(k), out) == 0
    return out.raw
