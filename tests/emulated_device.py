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
        self.streams = {}   # handle -> streamed witness
        self.graveyard = []
        self.hc = None
        self.hc_simt = None
        self.use_simt = False
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

    def b200_keccak256(self, data, n, out):
        from oracle.pyref import keccak256
        _wr(out, keccak256(bytes(data[:n]) if isinstance(data, (bytes, bytearray)) else _rd(data, n)))
        return 0

    def b200_host_alloc(self, nbytes, out_ptr):
        return self.b200_dev_alloc(nbytes, out_ptr)

    def b200_host_free(self, p):
        return self.b200_dev_free(p)

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

    def b200_axpy(self, fid, a, b, r, n, out):  # host pointers: the same thing here
        return self.b200_axpy_dev(fid, a, b, r, n, out, None)

    def b200_bind_top_multi_dev(self, fid, zs, k, n, r, stream):
        for j in range(k):
            rc = self.b200_bind_top_dev(fid, zs[j], n, r, stream)
            if rc:
                return rc
        return 0

    def b200_bind_top_dev(self, fid, z, n, r, stream):
        _wr(z, co.bind_top(fid, _rd(z, 32 * n), _rd(r, 32)))
        return 0

    def b200_sc_eval_dev(self, fid, form, A, B, C, length, eq_left, eq_right, shift, out, stream):
        if form == 11:  # SC_DOT: plain inner product (provider/ipa_pc.rs:102-108)
            P = FIELD_MODULUS[fid]
            self._put(fid, out, [sum(x * y for x, y in zip(self._ints(fid, A, length), self._ints(fid, B, length))) % P])
            return 0
        g = lambda p: _rd(p, 32 * length) if _addr(p) else None
        tab = lambda p: _rd(p, self._size(p)) if _addr(p) else None
        _wr(out, co.sc_eval(fid, form, g(A), g(B), g(C), tab(eq_left), tab(eq_right), shift))
        return 0

    def b200_sc_eval_sharded_dev(self, fid, form, A, B, C, length, eq_left, eq_right, shift, id_mul, id_add, out, stream):
        g = lambda p: _rd(p, 32 * length) if _addr(p) else None
        tab = lambda p: _rd(p, self._size(p)) if _addr(p) else None
        _wr(out, co.sc_eval(fid, form, g(A), g(B), g(C), tab(eq_left), tab(eq_right), shift, id_mul, id_add))
        return 0

    def b200_eq_table_dev(self, fid, r, ell, out, stream):
        _wr(out, co.eq_table(fid, _rd(r, 32 * ell)))
        return 0

    def b200_mle_eval_dev(self, fid, Z, ell, r, out, stream):
        _wr(out, co.mle_eval(fid, _rd(Z, 32 << ell), _rd(r, 32 * ell)))
        return 0

    def b200_mle_eval_multi_dev(self, fid, ptrs, k, ell, r, out, stream):
        rr = _rd(r, 32 * ell)
        _wr(out, b"".join(co.mle_eval(fid, _rd(ptrs[j], 32 << ell), rr) for j in range(k)))
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

    def b200_poly_eval_many_dev(self, fid, polys, lens, k, us, nu, evals, stream):
        for i in range(k):
            n = int(lens[i])
            out = ctypes.c_void_p(_addr(evals) + 32 * nu * i)
            if n == 0:
                _wr(out, bytes(32 * nu))
            else:
                self.b200_poly_eval_dev(fid, polys[i], n, us, nu, out, stream)
        return 0

    def b200_poly_div_dev(self, fid, f, n, u, out, stream):
        _wr(out, co.poly_div(fid, _rd(f, 32 * n), _rd(u, 32)))
        return 0

    # ---- sum-check loops with the transcript "on the device": the HOST BUILD of the round kernels -------
    # (tests/hostcheck, the same headers the GPU runs) strung together as csrc/capi_sumcheck.inc does
    def _hc(self):
        if self.hc is None:
            import os
            import subprocess
            here = os.path.dirname(os.path.abspath(__file__))
            src, so = os.path.join(here, "hostcheck", "hostcheck.cpp"), os.path.join(here, "hostcheck", "libhostcheck.so")
            csrc = os.path.join(here, "..", "nova_b200", "csrc")
            deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
            if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
            self.hc = ctypes.CDLL(so)
        return self.hc

    def _hc_simt(self):
        if self.hc_simt is None:
            import os
            import subprocess
            here = os.path.dirname(os.path.abspath(__file__))
            src, so = os.path.join(here, "hostcheck", "simt_check.cpp"), os.path.join(here, "hostcheck", "libhostcheck_simt.so")
            csrc = os.path.join(here, "..", "nova_b200", "csrc")
            deps = [src, os.path.join(here, "hostcheck", "simt_host.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
            if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
            self.hc_simt = ctypes.CDLL(so)
        return self.hc_simt

    def b200_sc_round_batched_dev(self, fid, desc, sums, state, pending, pending_len, la, ls, poly_out, r_out, stream):
        # use_simt: the real one-warp kernel on 32 host threads (tests/hostcheck/simt_host.h) instead of its body
        fn = self._hc_simt().hc_simt_sc_round_batched if self.use_simt else self._hc().hc_sc_round_batched
        rc = fn(fid, ctypes.c_void_p(_addr(desc)), ctypes.c_void_p(_addr(state)),
                                            ctypes.c_void_p(_addr(sums)), ctypes.c_void_p(_addr(pending)), int(pending_len),
                                            la, ls, ctypes.c_void_p(_addr(poly_out)), ctypes.c_void_p(_addr(r_out)))
        return 0 if rc == 0 else 1

    def _sumcheck_loop(self, fid, kind_of, claim, num_rounds, polys, tr, pending, pending_len, polys_out, r_out,
                       finals_out, ncoef, sums_of, taus=None):
        P = FIELD_MODULUS[fid]
        state = ctypes.create_string_buffer(_rd(claim, 32) + mont_bytes(P, 1) + _rd(tr, 72) + bytes(8), 144)
        length = 1 << num_rounds
        for j in range(num_rounds):
            res = ctypes.create_string_buffer(sums_of(j, length), 96)
            tau = tinv = None
            if taus is not None:
                tau = ctypes.create_string_buffer(mont_bytes(P, taus[j]), 32)
                tinv = ctypes.create_string_buffer(mont_bytes(P, pow(taus[j], -1, P) if taus[j] else 0), 32)
            rc = self._hc().hc_sc_round(fid, kind_of(j), state, res, tau, tinv, ctypes.c_void_p(_addr(pending) if j == 0 else 0),
                                        int(pending_len) if j == 0 else 0, ord("p"), ord("c"),
                                        ctypes.c_void_p(_addr(polys_out) + 32 * ncoef * j), ctypes.c_void_p(_addr(r_out) + 32 * j))
            assert rc == 0
            r = _rd(_addr(r_out) + 32 * j, 32)
            for Z in polys:
                _wr(Z, co.bind_top(fid, _rd(Z, 32 * length), r))
            length //= 2
        for k, Z in enumerate(polys):
            _wr(_addr(finals_out) + 32 * k, _rd(Z, 32))
        _wr(tr, state.raw[64:136])
        return 0

    def b200_sumcheck_quad_prod(self, fid, claim, num_rounds, A, B, tr, pending, pending_len, polys_out, r_out, finals_out):
        sums = lambda j, n: co.sc_eval(fid, 0, _rd(A, 32 * n), _rd(B, 32 * n))
        return self._sumcheck_loop(fid, lambda j: 0, claim, num_rounds, [A, B], tr, pending, pending_len, polys_out,
                                   r_out, finals_out, 2, sums)

    def b200_sumcheck_cubic3(self, fid, claim, taus_p, num_rounds, A, B, C, tr, pending, pending_len, polys_out, r_out,
                             finals_out):
        P = FIELD_MODULUS[fid]
        l = num_rounds
        raw = _rd(taus_p, 32 * l)
        taus = [from_mont_bytes(P, raw[32 * i:32 * i + 32]) for i in range(l)]
        fh, sh = l // 2, l - l // 2
        tab = lambda lo, hi: co.eq_table(fid, raw[32 * lo:32 * hi])
        left = [tab(fh - k, fh) for k in range(max(fh, 1))]
        right = [tab(l - k, l) for k in range(sh + 1)]

        def sums(j, n):
            rnd = j + 1
            L, R, shift = (left[fh - rnd], right[sh], sh) if rnd < fh else (None, right[l - rnd], 0)
            g = lambda Z: _rd(Z, 32 * n)
            out = co.sc_eval(fid, 4, g(A), g(B), g(C), L, R, shift)
            if taus[j] == 0:
                out += co.sc_eval(fid, 7, g(A), g(B), g(C), L, R, shift)
            return out
        return self._sumcheck_loop(fid, lambda j: 2 if taus[j] == 0 else 1, claim, num_rounds, [A, B, C], tr, pending,
                                   pending_len, polys_out, r_out, finals_out, 3, sums, taus)

    # ---- the whole batched sum-check (csrc/capi_sumcheck.inc b200_sumcheck_batched), restated: eq tables and sums by the
    # oracle, the round through the host build of the round kernel, binds by the oracle.  With `use_simt` the sums of
    # every round run through the REAL k_form_reduce_multi / k_form_final_multi and the last `tail_bits` variables
    # through the REAL k_scb_tail, on host threads (tests/hostcheck/simt_host.h). ----
    tail_bits = 2
    fused_round = True

    def b200_sumcheck_tail_bits(self, bits):
        old = self.tail_bits
        if bits >= 0:
            self.tail_bits = bits
        return old

    def b200_sumcheck_batched(self, fid, prog_ref, coeffs, claim, running, tr, pending, pending_len, polys_out, r_out,
                              finals_out):
        from nova_b200.ppsnark import SCB_MAX_CLAIMS, SCB_MAX_EQ, ScbDesc
        P = FIELD_MODULUS[fid]
        prog = prog_ref._obj
        l, nc, ne, nt = prog.num_rounds, prog.nclaims, prog.neq, prog.ntables
        fh, sh = l // 2, l - l // 2
        keep = []  # buffers the ctypes structs point into

        def buf(data: bytes):
            b = ctypes.create_string_buffer(data, max(len(data), 1))
            keep.append(b)
            return b
        eq = []
        for g in range(ne):
            raw = _rd(prog.taus[g], 32 * l)
            taus = [from_mont_bytes(P, raw[32 * i:32 * i + 32]) for i in range(l)]
            tab = lambda lo, hi: co.eq_table(fid, raw[32 * lo:32 * hi])
            left = buf(b"".join(tab(fh - k, fh) for k in range(max(fh, 1))))
            right = buf(b"".join(tab(l - k, l) for k in range(sh + 1)))
            tinv = buf(b"".join(mont_bytes(P, pow(t, -1, P) if t else 0) for t in taus))
            eq.append(dict(taus=taus, left=left, right=right, d_taus=buf(raw), d_tinv=tinv))
        state = buf(_rd(claim, 32) + bytes(32) + _rd(tr, 72) + bytes(8) + _rd(coeffs, 32 * nc) + bytes(32 * (SCB_MAX_CLAIMS - nc))
                    + _rd(running, 32 * nc) + bytes(32 * (SCB_MAX_CLAIMS - nc)) + mont_bytes(P, 1) * SCB_MAX_EQ)
        assert len(state.raw) == 1296
        tables = [prog.tables[t] for t in range(nt)]
        eq_of = [prog.eq_of[i] if prog.kind[i] >= 2 else -1 for i in range(nc)]

        def eq_ptrs(g, rnd):
            e = eq[g]
            if rnd < fh:
                return (ctypes.addressof(e["left"]) + 32 * ((1 << (fh - rnd)) - 1), 32 << (fh - rnd),
                        ctypes.addressof(e["right"]) + 32 * ((1 << sh) - 1), 32 << sh, sh)
            return 0, 0, ctypes.addressof(e["right"]) + 32 * ((1 << (l - rnd)) - 1), 32 << (l - rnd), 0
        length = 1 << l
        for j in range(l):
            if self.use_simt and l - j <= self.tail_bits:
                self._simt_tail(fid, prog, eq, eq_of, tables, j, state, pending, pending_len, polys_out, r_out)
                break
            rnd = j + 1
            plan = [(i, prog.form[i]) for i in range(nc)]
            plan += [(i, prog.form_m1[i]) for i in range(nc) if eq_of[i] >= 0 and eq[eq_of[i]]["taus"][j] == 0]
            d = ScbDesc()
            d.nclaims, d.neq = nc, ne
            for i in range(nc):
                d.kind[i], d.eq_of[i], d.slot[i], d.slot_m1[i] = prog.kind[i], eq_of[i], 3 * i, -1
            for k, (i, form) in enumerate(plan[nc:]):
                d.slot_m1[i] = 3 * (nc + k)
            for g in range(ne):
                d.tau[g] = ctypes.addressof(eq[g]["d_taus"]) + 32 * j
                d.tau_inv[g] = ctypes.addressof(eq[g]["d_tinv"]) + 32 * j
            sums = ctypes.create_string_buffer(96 * len(plan))
            if self.use_simt and self.fused_round:  # k_form_reduce_multi + k_sc_round_batched_fused, as the library does
                a = self._multi_args(fid, prog, plan, eq_of, eq_ptrs, rnd, tables, length)
                grid = max(1, min(3, (length // 2 + 255) // 256))
                rc = self._hc_simt().hc_simt_round_fused(
                    fid, ctypes.byref(a), grid, ctypes.byref(d), state, ctypes.c_void_p(_addr(pending) if j == 0 else 0),
                    int(pending_len) if j == 0 else 0, ord("p"), ord("c"), ctypes.c_void_p(_addr(polys_out) + 96 * j),
                    ctypes.c_void_p(_addr(r_out) + 32 * j))
                assert rc == 0
                r = _rd(_addr(r_out) + 32 * j, 32)
                for Z in tables:
                    _wr(Z, co.bind_top(fid, _rd(Z, 32 * length), r))
                length //= 2
                continue
            if self.use_simt:
                self._simt_multi(fid, prog, plan, eq_of, eq_ptrs, rnd, tables, length, sums)
            else:
                for k, (i, form) in enumerate(plan):
                    g = lambda c: _rd(tables[prog.tab[i][c]], 32 * length) if prog.tab[i][c] >= 0 else None
                    L = R = None
                    shift = 0
                    if eq_of[i] >= 0:
                        lp, ln, rp, rn, shift = eq_ptrs(eq_of[i], rnd)
                        L, R = (_rd(lp, ln) if lp else None), _rd(rp, rn)
                    res = co.sc_eval(fid, form, g(0), g(1), g(2), L, R, shift)
                    ctypes.memmove(ctypes.addressof(sums) + 96 * k, res, len(res))
            rc = self.b200_sc_round_batched_dev(fid, d, sums, state, pending if j == 0 else None, pending_len if j == 0 else 0,
                                                ord("p"), ord("c"), _addr(polys_out) + 96 * j, _addr(r_out) + 32 * j, None)
            assert rc == 0
            r = _rd(_addr(r_out) + 32 * j, 32)
            for Z in tables:
                _wr(Z, co.bind_top(fid, _rd(Z, 32 * length), r))
            length //= 2
        for t, Z in enumerate(tables):
            _wr(_addr(finals_out) + 32 * t, _rd(Z, 32))
        _wr(tr, state.raw[64:136])
        return 0

    def _simt_structs(self):
        """ctypes mirrors of multi_args (poly_kernels.cuh) and scb_tail_args (sumcheck_tail.cuh), checked by size"""
        from nova_b200.ppsnark import SCB_MAX_CLAIMS, SCB_MAX_EQ, ScbDesc
        if getattr(self, "_structs", None) is None:
            class MultiSum(ctypes.Structure):
                _fields_ = [("form", ctypes.c_int32), ("shift", ctypes.c_int32), ("A", ctypes.c_void_p), ("B", ctypes.c_void_p),
                            ("C", ctypes.c_void_p), ("eq_left", ctypes.c_void_p), ("eq_right", ctypes.c_void_p)]

            class MultiArgs(ctypes.Structure):
                _fields_ = [("n", ctypes.c_int32), ("h", ctypes.c_size_t), ("id_mul", ctypes.c_size_t), ("id_add", ctypes.c_size_t),
                            ("s", MultiSum * 32)]

            class TailEq(ctypes.Structure):
                _fields_ = [("left", ctypes.c_void_p), ("right", ctypes.c_void_p), ("taus", ctypes.c_void_p),
                            ("tau_inv", ctypes.c_void_p), ("tau_zero", ctypes.c_uint64)]

            class TailArgs(ctypes.Structure):
                _fields_ = [("d", ScbDesc), ("form", ctypes.c_int32 * SCB_MAX_CLAIMS), ("form_m1", ctypes.c_int32 * SCB_MAX_CLAIMS),
                            ("tab", (ctypes.c_int32 * 3) * SCB_MAX_CLAIMS), ("ntables", ctypes.c_int32),
                            ("num_rounds", ctypes.c_int32), ("first_round", ctypes.c_int32), ("tables", ctypes.c_void_p * 24),
                            ("eq", TailEq * SCB_MAX_EQ)]
            hc = self._hc_simt()
            assert ctypes.sizeof(MultiArgs) == hc.hc_simt_sizes(0) and ctypes.sizeof(MultiSum) == hc.hc_simt_sizes(1)
            assert ctypes.sizeof(TailArgs) == hc.hc_simt_sizes(2) and ctypes.sizeof(ScbDesc) == hc.hc_simt_sizes(3)
            self._structs = (MultiArgs, TailArgs)
        return self._structs

    def _simt_multi(self, fid, prog, plan, eq_of, eq_ptrs, rnd, tables, length, sums):
        a = self._multi_args(fid, prog, plan, eq_of, eq_ptrs, rnd, tables, length)
        grid = max(1, min(3, (length // 2 + 255) // 256))
        assert self._hc_simt().hc_simt_sc_reduce_multi(fid, ctypes.byref(a), grid, sums) == 0

    def _multi_args(self, fid, prog, plan, eq_of, eq_ptrs, rnd, tables, length):
        MultiArgs, _ = self._simt_structs()
        a = MultiArgs()
        a.n, a.h, a.id_mul, a.id_add = len(plan), length // 2, 1, 0
        for k, (i, form) in enumerate(plan):
            m = a.s[k]
            m.form = form
            m.A, m.B, m.C = (_addr(tables[prog.tab[i][c]]) if prog.tab[i][c] >= 0 else None for c in range(3))
            if eq_of[i] >= 0:
                lp, _, rp, _, shift = eq_ptrs(eq_of[i], rnd)
                m.eq_left, m.eq_right, m.shift = lp or None, rp, shift
        return a

    def _simt_tail(self, fid, prog, eq, eq_of, tables, j, state, pending, pending_len, polys_out, r_out):
        _, TailArgs = self._simt_structs()
        a = TailArgs()
        nc = prog.nclaims
        a.d.nclaims, a.d.neq = nc, prog.neq
        for i in range(nc):
            a.d.kind[i], a.d.eq_of[i], a.d.slot[i], a.d.slot_m1[i] = prog.kind[i], eq_of[i], 3 * i, -1
            a.form[i], a.form_m1[i] = prog.form[i], prog.form_m1[i] if eq_of[i] >= 0 else -1
            for c in range(3):
                a.tab[i][c] = prog.tab[i][c]
        a.ntables, a.num_rounds, a.first_round = prog.ntables, prog.num_rounds, j
        for t, Z in enumerate(tables):
            a.tables[t] = _addr(Z)
        for g, e in enumerate(eq):
            q = a.eq[g]
            q.left, q.right = ctypes.addressof(e["left"]), ctypes.addressof(e["right"])
            q.taus, q.tau_inv = ctypes.addressof(e["d_taus"]), ctypes.addressof(e["d_tinv"])
            q.tau_zero = sum(1 << k for k, t in enumerate(e["taus"]) if t == 0)
        sums = ctypes.create_string_buffer(96 * 32)
        rc = self._hc_simt().hc_simt_scb_tail(fid, ctypes.byref(a), state, sums, ctypes.c_void_p(_addr(pending) if j == 0 else 0),
                                              int(pending_len) if j == 0 else 0, ord("p"), ord("c"),
                                              ctypes.c_void_p(_addr(polys_out)), ctypes.c_void_p(_addr(r_out)))
        assert rc == 0

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

    # ---- inner-product argument kernels (include/nova_b200.h "inner-product argument") -------------
    def b200_fold_halves_dev(self, fid, v, n, x_lo, x_hi, out, stream):
        P = FIELD_MODULUS[fid]
        vs, lo, hi = self._ints(fid, v, n), self._ints(fid, x_lo, 1)[0], self._ints(fid, x_hi, 1)[0]
        self._put(fid, out, [(vs[i] * lo + vs[i + n // 2] * hi) % P for i in range(n // 2)])
        return 0

    def b200_ipa_scalars_dev(self, fid, a, w, n, nk, sL, sR, stream):
        P = FIELD_MODULUS[fid]
        h = nk // 2
        av, wv = self._ints(fid, a, nk), self._ints(fid, w, n)
        self._put(fid, sL, [av[j % h] * wv[j] % P if j & h else 0 for j in range(n)])
        self._put(fid, sR, [0 if j & h else av[(j % h) + h] * wv[j] % P for j in range(n)])
        return 0

    def b200_ipa_weights_dev(self, fid, w, n, nk, r, r_inv, stream):
        P = FIELD_MODULUS[fid]
        if nk == 0:
            self._put(fid, w, [1] * n)
            return 0
        rv, ri = self._ints(fid, r, 1)[0], self._ints(fid, r_inv, 1)[0]
        wv = self._ints(fid, w, n)
        self._put(fid, w, [wv[j] * (rv if j & (nk // 2) else ri) % P for j in range(n)])
        return 0

    # ---- host-pointer forms (same answers; "host" and "device" memory are the same thing here) ----
    def b200_sc_eval(self, fid, form, A, B, C, length, eq_left, eq_left_len, eq_right, eq_right_len, shift, out):
        return self.b200_sc_eval_dev(fid, form, A, B, C, length, eq_left, eq_right, shift, out, None)

    def b200_vec_add(self, fid, a, b, n, out):
        return self.b200_vec_add_dev(fid, a, b, n, out, None)

    def b200_cross_term(self, fid, az, bz, cz, e1, e2, u, n, t):
        return self.b200_cross_term_dev(fid, az, bz, cz, e1, e2, u, n, t, None)

    def b200_poly_eval(self, fid, f, n, us, nu, evals):
        return self.b200_poly_eval_dev(fid, f, n, us, nu, evals, None)

    def b200_poly_div(self, fid, f, n, u, out):
        return self.b200_poly_div_dev(fid, f, n, u, out, None)

    def b200_spmv_multi(self, handles, k, z1, z2, z_len, o1, o2):
        for j in range(k):
            if self.mats[handles[j]][5] != z_len:
                self.err = b"InvalidWitnessLength"
                return 1
            self.b200_spmv_dev(handles[j], z1, z2, o1[j], o2[j] if _addr(z2) else None, None)
        return 0

    # ---- commitment keys ----------------------------------------------------------------------
    def b200_msm_small(self, handle, off, scalars, elem_bytes, n, max_bits, out):
        curve_id, bases, _ = self.keys[handle]
        raw = _rd(scalars, elem_bytes * n)
        vals = [int.from_bytes(raw[elem_bytes * i:elem_bytes * (i + 1)], "little") for i in range(n)]
        _wr(out, self._jacobian(curve_id, co.msm_small(curve_id, vals, bases[64 * off:64 * (off + n)], max_bits if max_bits > 0 else -1)))
        return 0

    def b200_msm_indices(self, handle, idx, m, out):
        curve_id, bases, _ = self.keys[handle]
        _wr(out, self._jacobian(curve_id, co.batch_add(curve_id, bases, [int(idx[i]) for i in range(m)])))
        return 0

    def b200_commit_many_dev(self, handle, ptrs, lens, k, out, stream):
        for j in range(k):
            self.b200_commit_dev(handle, ptrs[j], lens[j], None, _addr(out) + 96 * j, stream)
        return 0

    def b200_msm_many_dev(self, handle, offsets, ptrs, lens, k, out, stream):
        for j in range(k):
            rc = self.b200_msm_dev(handle, offsets[j], ptrs[j], lens[j], _addr(out) + 96 * j, stream)
            if rc:
                return rc
        return 0

    def b200_ck_register(self, curve_id, bases, n, h, window_bits, out_handle):
        self.keys[self.next_handle] = (curve_id, _rd(bases, 64 * n), _rd(h, 64) if _addr(h) else None)
        out_handle._obj.value = self.next_handle
        self.next_handle += 1
        return 0

    def _first_invalid(self, curve_id, raw: bytes):
        """smallest index of a point with a non-canonical coordinate or off the curve (k_on_curve's contract)"""
        c = CURVES[curve_id]
        for i in range(len(raw) // 64):
            x, y = (int.from_bytes(raw[64 * i + k:64 * i + k + 32], "little") for k in (0, 32))
            if x >= c.p or y >= c.p or not c.on_curve(c.affine_from_bytes(raw[64 * i:64 * i + 64])):
                return i
        return None

    def b200_ck_validate(self, curve_id, bases, n, first_bad):
        bad = self._first_invalid(curve_id, _rd(bases, 64 * n))
        first_bad._obj.value = (1 << 64) - 1 if bad is None else bad
        return 0

    def b200_ck_register_checked(self, curve_id, bases, n, h, window_bits, out_handle, first_bad):
        raw = _rd(bases, 64 * n) + (_rd(h, 64) if _addr(h) else b"")
        bad = self._first_invalid(curve_id, raw)
        first_bad._obj.value = (1 << 64) - 1 if bad is None else bad
        out_handle._obj.value = 0
        if bad is not None:
            self.err = f"key point {bad} has a non-canonical coordinate or is not on the curve".encode()
            return 7
        return self.b200_ck_register(curve_id, bases, n, h, window_bits, out_handle)

    def b200_ck_setup_synthetic(self, curve_id, gen, k0, n, with_h, window_bits, out_handle):
        bases = co.gen_bases(curve_id, n + (1 if with_h else 0), int(k0))  # P_i = (k0 + i) G
        self.keys[self.next_handle] = (curve_id, bases[:64 * n], bases[64 * n:] if with_h else None)
        out_handle._obj.value = self.next_handle
        self.next_handle += 1
        return 0

    def b200_ck_setup_tau(self, curve_id, gen, tau_mont, n, window_bits, out_handle):
        from oracle import hyperkzg_ref as hk
        c = CURVES[curve_id]
        tau = from_mont_bytes(c.q, _rd(tau_mont, 32))
        self.keys[self.next_handle] = (curve_id, hk.setup_srs(curve_id, n, tau), None)  # ck[i] = [tau^i] G
        out_handle._obj.value = self.next_handle
        self.next_handle += 1
        return 0

    def b200_ck_export_bases(self, handle, offset, n, out):
        _, bases, _ = self.keys[handle]
        if 64 * (offset + n) > len(bases):
            return 5
        _wr(out, bases[64 * offset:64 * (offset + n)])
        return 0

    # ---- Poseidon RO (answered by oracle/poseidon_ref.py) ----------------------------------------------------
    def b200_poseidon_register(self, fid, arity, r_f, r_p, rc, mds, out_handle):
        from oracle import poseidon_ref as pr
        P = FIELD_MODULUS[fid]
        c = pr.cached_constants(P, arity)
        t = arity + 1
        got_rc = [from_mont_bytes(P, _rd(rc, 32 * (r_f + r_p) * t)[32 * i:32 * i + 32]) for i in range((r_f + r_p) * t)]
        got_m = [from_mont_bytes(P, _rd(mds, 32 * t * t)[32 * i:32 * i + 32]) for i in range(t * t)]
        if (r_f, r_p) != (c.r_f, c.r_p) or got_rc != c.rc or got_m != [x for row in c.m for x in row]:
            self.err = b"Poseidon constants differ from the oracle's"
            return 1
        if not hasattr(self, "poseidon"):
            self.poseidon = {}
        self.poseidon[self.next_handle] = (fid, arity)
        out_handle._obj.value = self.next_handle
        self.next_handle += 1
        return 0

    def b200_poseidon_ro_dev(self, handle, elems, n, num_bits, start_with_one, out, stream):
        from oracle import poseidon_ref as pr
        fid, arity = self.poseidon[handle]
        P = FIELD_MODULUS[fid]
        raw = _rd(elems, 32 * n)
        ro = pr.PoseidonRO(P, arity)
        for i in range(n):
            ro.absorb(from_mont_bytes(P, raw[32 * i:32 * i + 32]))
        c = ro.squeeze(num_bits, bool(start_with_one))
        h = ro.state[0]
        _wr(out, mont_bytes(P, h) + mont_bytes(P, c) + c.to_bytes(32, "little"))
        return 0

    def b200_poseidon_ro(self, handle, elems, n, num_bits, start_with_one, out):
        return self.b200_poseidon_ro_dev(handle, elems, n, num_bits, start_with_one, out, None)

    def b200_to_mont_dev(self, fid, src, n, dst, stream):
        P = FIELD_MODULUS[fid]
        raw = _rd(src, 32 * n)
        _wr(dst, b"".join(mont_bytes(P, int.from_bytes(raw[32 * i:32 * i + 32], "little")) for i in range(n)))
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

    def b200_msm_dev(self, handle, off, scalars, n, out, stream):
        curve_id, bases, _ = self.keys[handle]
        if off + n > len(bases) // 64:
            return 5
        _wr(out, self._jacobian(curve_id, co.msm(curve_id, _rd(scalars, 32 * n), bases[64 * off:64 * (off + n)])))
        return 0

    def b200_msm(self, handle, off, scalars, n, out):
        return self.b200_msm_dev(handle, off, scalars, n, out, None)

    def b200_jacobian_sum_dev(self, curve_id, pts, k, out, stream):
        c = CURVES[curve_id]
        acc = None
        raw = _rd(pts, 96 * k)
        for j in range(k):
            acc = c.add(acc, c.jacobian_from_bytes(raw[96 * j:96 * j + 96]))
        _wr(out, self._jacobian(curve_id, c.affine_bytes(acc)))
        return 0

    def b200_commit(self, handle, scalars, n, blind, out):  # host pointers: the same thing here
        return self.b200_commit_dev(handle, scalars, n, blind, out, None)

    # ---- streamed witness hand-off --------------------------------------------------------------
    def b200_witness_begin(self, ck, n, out_handle):
        buf = ctypes.create_string_buffer(max(32 * n, 1))
        self.streams[self.next_handle] = dict(ck=ck, n=n, buf=buf, filled=0, done=False)
        out_handle._obj.value = self.next_handle
        self.next_handle += 1
        return 0

    def b200_witness_append(self, h, scalars, count):
        w = self.streams[h]
        if w["done"] or w["filled"] + count > w["n"]:
            self.err = b"append overflows the witness"
            return 5
        ctypes.memmove(ctypes.addressof(w["buf"]) + 32 * w["filled"], _rd(scalars, 32 * count), 32 * count)
        w["filled"] += count
        return 0

    def b200_witness_finish(self, h, r, out, d_witness):
        w = self.streams[h]
        if w["done"]:
            self.err = b"witness stream already finished"
            return 1
        w["done"] = True
        rc = self.b200_commit_dev(w["ck"], w["buf"], w["n"], r, out, None)
        if d_witness is not None:
            d_witness._obj.value = ctypes.addressof(w["buf"])
        return rc

    def b200_witness_reset(self, h):
        w = self.streams[h]
        ctypes.memset(w["buf"], 0, len(w["buf"]))
        w["filled"], w["done"] = 0, False
        return 0

    def b200_witness_release(self, h):
        self.streams.pop(h, None)
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
