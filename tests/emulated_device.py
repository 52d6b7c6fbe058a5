"""A CPU stand-in for libnova_b200.so, for testing the HOST LOGIC of the Python mirror without a GPU.

TEST INFRASTRUCTURE ONLY.  "Device memory" is host memory, and every `*_dev` entry point the mirror's
composed provers call is answered by the C oracle on the bytes behind the pointers.  Installing it
(`install()`, undone by `uninstall()`) makes `nova_b200.native.lib()` return this object, so the very
same mirror code that drives the GPU (nova_b200/snark.py, spartan.py, ...) runs here and its glue --
buffer sizes, offsets, argument order, transcript labels, claim bookkeeping -- is checked against the
independent Python restatements.  It proves nothing about the CUDA kernels; the `-m gpu` tests do that.

Only the subset the CPU tests need is implemented; anything else raises AttributeError loudly.
"""
import ctypes

from oracle import coracle as co
from oracle.pyref import CURVES, FIELD_MODULUS, from_mont_bytes, mont_bytes


def _addr(x) -> int:
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if isinstance(x, ctypes.c_void_p):
        return x.value or 0
    if hasattr(x, "_obj"):  # ctypes.byref(obj)
        return ctypes.addressof(x._obj)
    return ctypes.addressof(x)


def _rd(x, nbytes: int) -> bytes:
    return ctypes.string_at(_addr(x), nbytes) if nbytes else b""


def _wr(x, data: bytes):
    if data:
        ctypes.memmove(_addr(x), data, len(data))


class EmulatedDevice:
    def __init__(self):
        self.allocs = {}    # address -> buffer (keeps it alive), gives sizes of whole allocations
        self.mats = {}      # handle -> (fid, data, indices, indptr, rows, cols)
        self.keys = {}      # handle -> (curve, bases, h or None)
        self.graveyard = []
        self.next_handle = 1
        self.err = b""

    # ---- library / memory ---------------------------------------------------------------------
    def b200_init(self, device):
        return 0

    def b200_last_error(self):
        return self.err

    def b200_sync(self):
        return 0

    def b200_dev_alloc(self, nbytes, out_ptr):
        buf = ctypes.create_string_buffer(max(int(nbytes), 1))
        a = ctypes.addressof(buf)
        self.allocs[a] = buf
        out_ptr._obj.value = a
        return 0

    def b200_dev_free(self, p):
        buf = self.allocs.pop(_addr(p), None)
        if buf is not None:  # poison: a later read through a dangling pointer yields garbage, not stale data
            ctypes.memset(buf, 0xEE, len(buf))
            self.graveyard.append(buf)  # keep the pages mapped so that such a read cannot crash the test run
        return 0

    def b200_memcpy_h2d(self, d, h, n):
        _wr(d, _rd(h, n))
        return 0

    def b200_memcpy_d2h(self, h, d, n):
        _wr(h, _rd(d, n))
        return 0

    def b200_memcpy_d2d(self, dst, src, n, stream):
        _wr(dst, _rd(src, n))
        return 0

    def b200_memset_dev(self, d, byte, n, stream):
        _wr(d, bytes([byte]) * n)
        return 0

    def _size(self, p) -> int:
        return len(self.allocs[_addr(p)])

    # ---- field vectors ------------------------------------------------------------------------
    def b200_axpy_dev(self, fid, a, b, r, n, out, stream):
        _wr(out, co.axpy(fid, _rd(a, 32 * n), _rd(b, 32 * n), _rd(r, 32)))
        return 0

    def b200_bind_top_dev(self, fid, z, n, r, stream):
        _wr(z, co.bind_top(fid, _rd(z, 32 * n), _rd(r, 32)))
        return 0

    def b200_sc_eval_dev(self, fid, form, A, B, C, length, eq_left, eq_right, shift, out, stream):
        g = lambda p: _rd(p, 32 * length) if _addr(p) else None
        tab = lambda p: _rd(p, self._size(p)) if _addr(p) else None
        _wr(out, co.sc_eval(fid, form, g(A), g(B), g(C), tab(eq_left), tab(eq_right), shift))
        return 0

    def b200_eq_table_dev(self, fid, r, ell, out, stream):
        _wr(out, co.eq_table(fid, _rd(r, 32 * ell)))
        return 0

    def b200_mle_eval_dev(self, fid, Z, ell, r, out, stream):
        _wr(out, co.mle_eval(fid, _rd(Z, 32 << ell), _rd(r, 32 * ell)))
        return 0

    def b200_rlc_dev(self, fid, ptrs, lens, k, coeffs, n, out, stream):
        polys = [_rd(ptrs[i], 32 * lens[i]) for i in range(k)]
        _wr(out, co.rlc(fid, polys, _rd(coeffs, 32 * k), n))
        return 0

    def b200_kzg_fold_dev(self, fid, p_, n, x, out, stream):
        _wr(out, co.kzg_fold(fid, _rd(p_, 32 * n), _rd(x, 32)))
        return 0

    def b200_poly_eval_dev(self, fid, f, n, us, nu, evals, stream):
        _wr(evals, co.poly_eval(fid, _rd(f, 32 * n), _rd(us, 32 * nu)))
        return 0

    def b200_poly_div_dev(self, fid, f, n, u, out, stream):
        _wr(out, co.poly_div(fid, _rd(f, 32 * n), _rd(u, 32)))
        return 0

    # ---- sparse matrices ----------------------------------------------------------------------
    def b200_spmv_register(self, fid, data, indices, indptr, rows, cols, out_handle):
        ip = [int(indptr[i]) for i in range(rows + 1)]
        nnz = ip[-1]
        self.mats[self.next_handle] = (fid, _rd(data, 32 * nnz), [int(indices[i]) for i in range(nnz)], ip, rows, cols)
        out_handle._obj.value = self.next_handle
        self.next_handle += 1
        return 0

    def b200_spmv_release(self, handle):
        self.mats.pop(handle, None)
        return 0

    def b200_spmv_dev(self, handle, z1, z2, o1, o2, stream):
        fid, data, idx, ip, rows, cols = self.mats[handle]
        _wr(o1, co.spmv(fid, data, idx, ip, _rd(z1, 32 * cols)))
        if _addr(z2):
            _wr(o2, co.spmv(fid, data, idx, ip, _rd(z2, 32 * cols)))
        return 0

    def b200_spmv_t_dev(self, handle, rx, out_len, out, stream):
        fid, data, idx, ip, rows, cols = self.mats[handle]
        _wr(out, co.spmv_t(fid, data, idx, ip, _rd(rx, 32 * rows), out_len))
        return 0

    def b200_vec_add_dev(self, fid, a, b, n, out, stream):
        _wr(out, co.vec_add(fid, _rd(a, 32 * n), _rd(b, 32 * n)))
        return 0

    def b200_cross_term_dev(self, fid, az, bz, cz, e1, e2, u, n, t, stream):
        g = lambda p: _rd(p, 32 * n)
        _wr(t, co.cross_term(fid, g(az), g(bz), g(cz), g(e1), g(e2) if _addr(e2) else None, _rd(u, 32)))
        return 0

    def _ints(self, fid, p_, n):
        P = FIELD_MODULUS[fid]
        raw = _rd(p_, 32 * n)
        return [from_mont_bytes(P, raw[32 * i:32 * i + 32]) for i in range(n)]

    def _put(self, fid, p_, xs):
        P = FIELD_MODULUS[fid]
        _wr(p_, b"".join(mont_bytes(P, x) for x in xs))

    def b200_vec_mul_dev(self, fid, a, b, n, out, stream):
        P = FIELD_MODULUS[fid]
        self._put(fid, out, [x * y % P for x, y in zip(self._ints(fid, a, n), self._ints(fid, b, n))])
        return 0

    def b200_logup_hash_dev(self, fid, val, addr, gamma, r, n, out, stream):
        """out[i] = val[i] * gamma + addr[i] + r, addr == NULL meaning the cell's own index (ppsnark.rs:386-435)."""
        P = FIELD_MODULUS[fid]
        g, rr = self._ints(fid, gamma, 1)[0], self._ints(fid, r, 1)[0]
        ad = self._ints(fid, addr, n) if _addr(addr) else list(range(n))
        self._put(fid, out, [(v * g + a + rr) % P for v, a in zip(self._ints(fid, val, n), ad)])
        return 0

    def b200_batch_invert_dev(self, fid, inp, n, out, zero_flag, stream):
        res = co.batch_invert(fid, _rd(inp, 32 * n))
        _wr(zero_flag, (1 if res is None else 0).to_bytes(4, "little"))
        if res is not None:
            _wr(out, res)
        return 0

    def b200_gather_dev(self, table, idx, n, out, stream):
        ix = (ctypes.c_uint32 * n).from_address(_addr(idx))
        base = _addr(table)
        _wr(out, b"".join(ctypes.string_at(base + 32 * int(ix[i]), 32) for i in range(n)))
        return 0

    # ---- commitment keys ----------------------------------------------------------------------
    def b200_commit_many_dev(self, handle, ptrs, lens, k, out, stream):
        for j in range(k):
            self.b200_commit_dev(handle, ptrs[j], lens[j], None, _addr(out) + 96 * j, stream)
        return 0

    def b200_ck_register(self, curve_id, bases, n, h, window_bits, out_handle):
        self.keys[self.next_handle] = (curve_id, _rd(bases, 64 * n), _rd(h, 64) if _addr(h) else None)
        out_handle._obj.value = self.next_handle
        self.next_handle += 1
        return 0

    def b200_ck_release(self, handle):
        self.keys.pop(handle, None)
        return 0

    def _jacobian(self, curve_id, affine: bytes) -> bytes:
        z = bytes(32) if affine == bytes(64) else mont_bytes(CURVES[curve_id].p, 1)
        return affine + z  # (x, y, 1) or the identity (z = 0)

    def b200_commit_dev(self, handle, scalars, n, blind, out, stream):
        curve_id, bases, h = self.keys[handle]
        sc, bs = _rd(scalars, 32 * n), bases[:64 * n]
        if _addr(blind):
            sc, bs = sc + _rd(blind, 32), bs + h
        _wr(out, self._jacobian(curve_id, co.msm(curve_id, sc, bs)))
        return 0

    # ---- group --------------------------------------------------------------------------------
    def b200_msm_adhoc(self, curve_id, bases, scalars, n, out):
        _wr(out, self._jacobian(curve_id, co.msm(curve_id, _rd(scalars, 32 * n), _rd(bases, 64 * n))))
        return 0


_saved = []


def install() -> EmulatedDevice:
    import nova_b200.native as native
    dev = EmulatedDevice()
    _saved.append(native._LIB)
    native._LIB = dev
    return dev


def uninstall():
    import nova_b200.native as native
    from nova_b200 import spartan
    spartan._SMALL.clear()  # cached scratch vectors point into the emulated memory
    native._LIB = _saved.pop() if _saved else None
