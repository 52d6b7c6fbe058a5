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


# ---- sum-check / MLE / HyperKZG / SpMV -----------------------------------------------------------
SC_NOUT = {0: 2, 1: 2, 2: 2, 3: 3, 4: 2, 5: 2, 6: 1, 7: 1, 8: 1, 9: 1, 10: 1}


def sc_eval(fid, form, A, B=None, C=None, eq_left=None, eq_right=None, shift=0, id_mul=1, id_add=0) -> bytes:
    n = len(A) // 32
    out = ctypes.create_string_buffer(96)
    rc = lib().orc_sc_eval_sharded(fid, form, _buf(A), _buf(B) if B else None, _buf(C) if C else None,
                                   ctypes.c_size_t(n), _buf(eq_left) if eq_left else None,
                                   _buf(eq_right) if eq_right else None, shift, ctypes.c_size_t(id_mul),
                                   ctypes.c_size_t(id_add), out)
    assert rc == 0
    return out.raw[: 32 * SC_NOUT[form]]


def eq_table(fid, r: bytes) -> bytes:
    ell = len(r) // 32
    out = ctypes.create_string_buffer(32 << ell)
    assert lib().orc_eq_table(fid, _buf(r), ell, out) == 0
    return out.raw


def mle_eval(fid, Z: bytes, r: bytes) -> bytes:
    ell = len(r) // 32
    assert len(Z) == 32 << ell
    out = ctypes.create_string_buffer(32)
    assert lib().orc_mle_eval(fid, _buf(Z), ell, _buf(r), out) == 0
    return out.raw


def batch_invert(fid, v: bytes):
    n = len(v) // 32
    out = ctypes.create_string_buffer(max(32 * n, 1))
    rc = lib().orc_batch_invert(fid, _buf(v), ctypes.c_size_t(n), out)
    return None if rc == 2 else out.raw[: 32 * n]


def rlc(fid, polys, coeffs: bytes, n: int) -> bytes:
    k = len(polys)
    bufs = [_buf(p) for p in polys]
    ptrs = (ctypes.c_void_p * max(k, 1))(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
    lens = (ctypes.c_size_t * max(k, 1))(*[len(p) // 32 for p in polys])
    out = ctypes.create_string_buffer(max(32 * n, 1))
    assert lib().orc_rlc(fid, ptrs, lens, ctypes.c_size_t(k), _buf(coeffs), ctypes.c_size_t(n), out) == 0
    return out.raw[: 32 * n]


def kzg_fold(fid, p: bytes, x: bytes) -> bytes:
    n = len(p) // 32
    out = ctypes.create_string_buffer(max(16 * n, 1))
    assert lib().orc_kzg_fold(fid, _buf(p), ctypes.c_size_t(n), _buf(x), out) == 0
    return out.raw[: 16 * n]


def poly_eval(fid, f: bytes, us: bytes) -> bytes:
    n, nu = len(f) // 32, len(us) // 32
    out = ctypes.create_string_buffer(32 * nu)
    assert lib().orc_poly_eval(fid, _buf(f), ctypes.c_size_t(n), _buf(us), ctypes.c_size_t(nu), out) == 0
    return out.raw


def poly_div(fid, f: bytes, u: bytes) -> bytes:
    n = len(f) // 32
    out = ctypes.create_string_buffer(max(32 * (n - 1), 1))
    assert lib().orc_poly_div(fid, _buf(f), ctypes.c_size_t(n), _buf(u), out) == 0
    return out.raw[: 32 * (n - 1)]


def spmv(fid, data: bytes, indices, indptr, z: bytes) -> bytes:
    rows = len(indptr) - 1
    ia = (ctypes.c_uint64 * max(len(indices), 1))(*indices)
    ip = (ctypes.c_uint64 * len(indptr))(*indptr)
    out = ctypes.create_string_buffer(max(32 * rows, 1))
    assert lib().orc_spmv(fid, _buf(data), ia, ip, ctypes.c_size_t(rows), _buf(z), out) == 0
    return out.raw[: 32 * rows]


def spmv_t(fid, data: bytes, indices, indptr, rx: bytes, out_len: int) -> bytes:
    rows = len(indptr) - 1
    ia = (ctypes.c_uint64 * max(len(indices), 1))(*indices)
    ip = (ctypes.c_uint64 * len(indptr))(*indptr)
    out = ctypes.create_string_buffer(max(32 * out_len, 1))
    assert lib().orc_spmv_t(fid, _buf(data), ia, ip, ctypes.c_size_t(rows), _buf(rx), ctypes.c_size_t(out_len), out) == 0
    return out.raw[: 32 * out_len]
