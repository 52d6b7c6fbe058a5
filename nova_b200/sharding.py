"""Multi-GPU sharding of the MSM path: one process per GPU, the (scalar, base)
pairs split by INDEX RANGE, each rank reduces its slice to one point, and the partial points are
exchanged with a single all-gather (NCCL has no group-law reduction, so all-reduce is realised as
all-gather + local add of `world` points).  The key slice a rank owns is fixed for the life of the
key, so only scalars move.

The functions here are backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank`; ranges tile [0, n) and differ by <= 1."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def all_gather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """partial: uint8[96] Jacobian point of this rank -> uint8[world*96] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return partial.clone()
    out = torch.empty(96 * world, dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out


# ---------------------------------------------------------------------------------------------
# sum-check sharding: cyclic distribution of the polynomial tables
# ---------------------------------------------------------------------------------------------
def cyclic_shard(vec: bytes, rank: int, world: int, elem: int = 32) -> bytes:
    """Entries rank, rank + world, ... of a table of `elem`-byte elements."""
    n = len(vec) // elem
    return b"".join(vec[elem * i:elem * i + elem] for i in range(rank, n, world))


def all_gather_bytes(b: bytes, group=None) -> list:
    """Every rank's equally sized byte string (a few field elements), in rank order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [b]
    if dist.get_backend(group) == "nccl":  # NCCL moves device tensors only: stage the few bytes through the GPU
        if not b:
            return [b""] * world
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        out = torch.empty(len(b) * world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(out, t, group=group)
        raw = out.cpu().numpy().tobytes()
        return [raw[len(b) * k:len(b) * (k + 1)] for k in range(world)]
    t = torch.frombuffer(bytearray(b), dtype=torch.uint8)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [bytes(x.numpy().tobytes()) for x in out]


def sharded_prove_cubic_with_three_inputs(engine, p: int, claim: int, taus: list, A, B, C, transcript,
                                          rank: int, world: int, group=None, gather=None):
    """`SumcheckProof::prove_cubic_with_three_inputs` (src/spartan/sumcheck.rs:446-507) with the three
    tables sharded cyclically over `world` ranks.

    `engine` supplies the local O(N/world) work on its own table handles:
        engine.upload(bytes) -> handle ; engine.download(handle, n_elems) -> bytes
        engine.eq_tables(taus) -> object with .tables(round) -> (left|None, right, shift)
        engine.sc_eval(form, A, B, C, local_len, left, right, shift, id_mul, id_add) -> list[int]
        engine.bind(handle, local_len, r) -> None            (in place, top variable)
    Per round the ranks exchange 2 (or 1 more on the tau = 0 fall-back) field elements; binding needs
    no exchange because (i, i + len/2) are co-resident under the cyclic layout.  When one element per
    rank is left, the `world` values are all-gathered and the last log2(world) rounds run replicated.
    Every rank returns the same (compressed polys, challenges, final evaluations).
    `gather(bytes) -> list of every rank's bytes` replaces the default host all-gather (e.g. NcclComm.gather_bytes
    when the process group is NCCL, which takes no CPU tensors).
    """
    if gather is None:
        gather = lambda b: all_gather_bytes(b, group)
    from .spartan import UniPoly  # O(1) host algebra shared with the single-GPU prover

    l = len(taus)
    n = 1 << l
    assert world & (world - 1) == 0 and world <= n
    local_len = n // world
    eq = engine.eq_tables(taus)
    eval_eq_left = 1
    tau_consts = [((1 - t) % p, (2 * t - 1) % p, (2 - 3 * t) % p) for t in taus]
    rs, polys = [], []
    id_mul, id_add = world, rank
    for rnd in range(1, l + 1):
        if local_len == 1 and id_mul > 1:
            # global length == world: replicate the tail on every rank
            parts = [gather(engine.download(h, 1)) for h in (A, B, C)]
            A, B, C = (engine.upload(b"".join(pp)) for pp in parts)
            local_len, id_mul, id_add = world, 1, 0
        left, right, shift = eq.tables(rnd)

        def gsum(form):
            vals = engine.sc_eval(form, A, B, C, local_len, left, right, shift, id_mul, id_add)
            if id_mul == 1:
                return vals
            raw = b"".join(int(v).to_bytes(32, "little") for v in vals)
            allv = gather(raw)
            return [sum(int.from_bytes(x[32 * k:32 * k + 32], "little") for x in allv) % p
                    for k in range(len(vals))]

        t0, tinf = gsum(4)  # SC_EQ_CUBIC3
        e0c, slope, em1 = tau_consts[rnd - 1]
        l1p = (e0c + slope) * eval_eq_left % p
        if l1p != 0:  # derive_from_claim_deg2, sumcheck.rs:680-715
            s0 = e0c * eval_eq_left * t0 % p
            t1 = (claim - s0) * pow(l1p, -1, p) % p
            lead = slope * eval_eq_left * tinf % p
            sm1 = em1 * eval_eq_left * ((2 * tinf + 2 * t0 - t1) % p) % p
        else:  # tau = 0 fall-back, sumcheck.rs:1082-1130
            (tm1,) = gsum(7)  # SC_EQ_CUBIC3_M1
            s0 = e0c * eval_eq_left * t0 % p
            lead = slope * eval_eq_left * tinf % p
            sm1 = em1 * eval_eq_left * tm1 % p
        poly = UniPoly.from_evals_deg3(p, [s0, (claim - s0) % p, lead, sm1])
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        polys.append(poly.compress())
        claim = poly.evaluate(r)
        for h in (A, B, C):
            engine.bind(h, local_len, r)
        local_len //= 2
        tau = taus[rnd - 1]
        eval_eq_left = eval_eq_left * (1 - tau - r + 2 * r * tau) % p  # EqSumCheckInstance::bound
    finals = [int.from_bytes(engine.download_canonical(h), "little") for h in (A, B, C)]
    return polys, rs, finals


# ---------------------------------------------------------------------------------------------
# the other pieces of SURVEY.md §8e: SpMV, the folding step, Horner evaluation, division by (X - u)
# ---------------------------------------------------------------------------------------------
def all_gather_var(b: bytes, group=None) -> list:
    """Every rank's byte string, of possibly different lengths (index-range slices differ by one element)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [b]
    lens = [int.from_bytes(x, "little") for x in all_gather_bytes(len(b).to_bytes(8, "little"), group)]
    m = max(lens)
    padded = all_gather_bytes(b + bytes(m - len(b)), group) if m else [b""] * world
    return [x[:n] for x, n in zip(padded, lens)]


class DeviceEngine:
    """The per-rank engine of the functions below on a GPU: thin adapters over the host-pointer mirror calls
    (every rank drives its own device).  The CPU tests pass an oracle-backed object with the same methods."""

    def __init__(self, fid: int, curve=None):
        from . import provider, spartan
        self.fid, self.curve, self._sp, self._pv = fid, curve, spartan, provider

    def matrix(self, data: bytes, indices, indptr, cols: int):
        return self._sp.SparseMatrix(self.fid, data, indices, indptr, cols)

    def spmv(self, mat, z: bytes) -> bytes:
        return mat.multiply_vec(z)

    def vec_add(self, a: bytes, b: bytes) -> bytes:
        return self._pv.vec_add(self.fid, a, b)

    def cross_term(self, az, bz, cz, e1, u: bytes, e2=None) -> bytes:
        return self._pv.cross_term(self.fid, az, bz, cz, e1, u, e2)

    def poly_eval(self, f: bytes, u: bytes) -> bytes:
        return self._sp.poly_eval(self.fid, f, u)

    def poly_div(self, f: bytes, u: bytes) -> bytes:
        return self._sp.poly_div(self.fid, f, u)

    def axpy(self, a: bytes, b: bytes, r: bytes) -> bytes:
        return self._pv.fold_witness(self.fid, a, b, r)

    def msm_partial(self, ck, base_offset: int, scalars: bytes) -> bytes:
        """sum_i scalars[i] * ck[base_offset + i] as a 96-byte Jacobian point (identity for no scalars)"""
        import ctypes

        from .native import check, lib
        out = ctypes.create_string_buffer(96)
        check(lib().b200_msm(ck.handle, base_offset, self._pv._cbuf(scalars), len(scalars) // 32, out))
        return out.raw

    def point_sum(self, parts: list):
        from .native import check, lib
        d, out = self._sp.DeviceVec.from_bytes(b"".join(parts)), self._sp.DeviceVec(96)
        check(lib().b200_jacobian_sum_dev(int(self.curve), d.ptr, len(parts), out.ptr, None))
        raw = out.to_bytes()
        d.free()
        out.free()
        return self._pv._jac_to_affine(self._pv.Curve(self.curve), raw)


def sharded_multiply_vec(engine, mats_rows, z_local: bytes, group=None):
    """R1CSShape::multiply_vec (src/r1cs/mod.rs:407-431) with the ROWS of A, B, C split by index range and z
    split the same way: z is all-gathered once (n x 32 B), every rank multiplies its own rows.
    mats_rows: this rank's (A, B, C) row slices as engine matrices.  -> (z, (Az, Bz, Cz) local rows)."""
    z = b"".join(all_gather_var(z_local, group))
    return z, tuple(engine.spmv(M, z) for M in mats_rows)


def sharded_cross_term(engine, mats_rows, z1_local: bytes, z2_local: bytes, e1_rows: bytes, u_sum: bytes,
                       e2_rows: bytes | None = None, group=None) -> bytes:
    """The T of commit_T / commit_T_relaxed (src/r1cs/mod.rs:578-664) on row slices: Z1 + Z2 locally, one
    all-gather of Z, local SpMV rows, local T rows.  T is born on the rank that owns the same index range of
    the commitment key, so its MSM needs only the 96-byte partial exchange (all_gather_partials)."""
    _, (az, bz, cz) = sharded_multiply_vec(engine, mats_rows, engine.vec_add(z1_local, z2_local), group)
    return engine.cross_term(az, bz, cz, e1_rows, u_sum, e2_rows)


def sharded_fold_step(engine, p: int, ck, mats_rows, rows: tuple, wrange: tuple, z1_local: bytes, z2_local: bytes,
                      W1_local: bytes, E1_rows: bytes, W2_local: bytes, u_sum: bytes, challenge, group=None):
    """The folding half of `prove_step` (nova/mod.rs:456-564: commit(W2), `commit_T` r1cs/mod.rs:578-627,
    `RelaxedR1CSWitness::fold` :1044-1069) with every vector split by index range over the ranks:
    W by [wlo, whi) = `wrange`, the rows of A, B, C, E and T by [rlo, rhi) = `rows`, z like the matrices' columns
    (`z*_local`).  `ck` is a key that holds ck[0 .. max(num_vars, num_cons)) -- the full key, or any key whose
    indices line up with the global ones.
      1. partial commitments of W2 (over ck[wlo..whi)) and, after the one all-gather of Z and the local rows of
         T, of T (over ck[rlo..rhi)): ONE all-gather of 2 x 96 bytes, local sums -> comm_W2, comm_T
      2. r = challenge(comm_W2, comm_T) -- the random-oracle step, replicated on every rank
      3. W = W1 + r W2 and E = E1 + r T on the local slices, no exchange.
    `engine` adds two methods to the ones of `sharded_cross_term`: msm_partial(ck, base_offset, scalars) -> 96-byte
    Jacobian partial, point_sum(list of 96-byte points) -> affine, and axpy(a, b, r) -> a + r b.
    -> (comm_W2, comm_T, r, W_local, E_rows, T_rows)"""
    rlo, rhi = rows
    wlo, whi = wrange
    T_rows = sharded_cross_term(engine, mats_rows, z1_local, z2_local, E1_rows, u_sum, None, group)
    part = engine.msm_partial(ck, wlo, W2_local) + engine.msm_partial(ck, rlo, T_rows)
    allp = all_gather_bytes(part, group)
    comm_W2 = engine.point_sum([a[:96] for a in allp])
    comm_T = engine.point_sum([a[96:] for a in allp])
    r = challenge(comm_W2, comm_T)
    rm = _mont(p, r)
    return comm_W2, comm_T, r, engine.axpy(W1_local, W2_local, rm), engine.axpy(E1_rows, T_rows, rm), T_rows


def _mont(p: int, x: int) -> bytes:
    return ((x % p) * (1 << 256) % p).to_bytes(32, "little")


def _unmont(p: int, b: bytes) -> int:
    return int.from_bytes(b, "little") * pow(1 << 256, -1, p) % p


def sharded_poly_eval(engine, p: int, f_local: bytes, lo: int, u: int, group=None) -> int:
    """f(u) for coefficients split by index range (the 3-point evaluations of hyperkzg.rs:1048-1056):
    every rank evaluates its slice as a polynomial in X, scales by u^lo, and the `world` values are
    all-gathered and added."""
    v = _unmont(p, engine.poly_eval(f_local, _mont(p, u))) if f_local else 0
    part = v * pow(u, lo, p) % p
    return sum(int.from_bytes(x, "little") for x in all_gather_bytes(part.to_bytes(32, "little"), group)) % p


def sharded_poly_div(engine, p: int, f_local: bytes, lo: int, hi: int, n: int, u: int, group=None) -> bytes:
    """The quotient of f by (X - u) (hyperkzg.rs:961-999 `div_by_monomial`: h_{n-2} = f_{n-1},
    h_{i-1} = f_i + u h_i) with f split by index range: rank g returns h[lo .. hi) (the last rank one entry
    less, n - 1 in total).  With H_i = sum_{j >= i} f_j u^(j-i): every rank all-gathers its slice value
    V_g = H restricted to its slice, the carry into rank g is C_g = sum_{k > g} V_k u^(lo_k - hi_g), and
    dividing the slice extended by the coefficient C_g yields exactly h[lo .. hi)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    um = _mont(p, u)
    v = _unmont(p, engine.poly_eval(f_local, um)) if f_local else 0
    meta = all_gather_bytes(v.to_bytes(32, "little") + lo.to_bytes(8, "little") + hi.to_bytes(8, "little"), group)
    rank = dist.get_rank(group) if world > 1 else 0
    carry = 0
    for k in range(rank + 1, world):
        vk, lok = int.from_bytes(meta[k][:32], "little"), int.from_bytes(meta[k][32:40], "little")
        carry = (carry + vk * pow(u, lok - hi, p)) % p
    if not f_local:
        return b""
    q = engine.poly_div(f_local + _mont(p, carry), um)  # (hi - lo) entries: h[lo .. hi)
    return q[:32 * (hi - lo - 1)] if hi == n else q


# ---------------------------------------------------------------------------------------------
# HyperKZG EvaluationEngine::prove over `world` GPUs (SURVEY.md §8d config C4, §8e rows "HyperKZG fold",
# "Horner evals / div-by-monomial", "MSM")
# ---------------------------------------------------------------------------------------------
class HostStagedComm:
    """Collectives on device buffers staged through the host (works with any torch.distributed backend that
    takes CPU tensors, i.e. gloo): D2H of the local piece, all-gather, H2D of the result.  The four small
    exchanges of the prover move a few hundred bytes; only the fold-chain gather is bulk data."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def gather_bytes(self, b: bytes) -> list:
        return all_gather_bytes(b, self.group)

    def gather_dev(self, vec, nbytes: int):
        """every rank's `nbytes` of `vec`, concatenated in rank order, as a new device vector"""
        from .spartan import DeviceVec
        return DeviceVec.from_bytes(b"".join(all_gather_bytes(vec.to_bytes(nbytes), self.group)))


class _CudaView:
    """zero-copy view of a raw device allocation for torch (CUDA array interface, version 2)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class NcclComm:
    """The same two collectives on NCCL, device to device: the library's allocations are viewed as torch tensors
    (no copy) and all-gathered over NVLink; the small host-side exchanges go through a CUDA staging tensor.
    The library works on its own stream, torch/NCCL on torch's: both sides are synchronised around every
    collective (six per proof).  Verified on 2 / 4 / 8 B200s (proofs accepted by the restated verifier and equal to the
    single-GPU proof: profiles/r02c, r02e, r02k); tests/test_zz_new_paths_gpu.py::test_sharded_hyperkzg_nccl needs two GPUs."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def gather_bytes(self, b: bytes) -> list:
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        out = torch.empty(len(b) * self.world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(out, t, group=self.group)
        raw = out.cpu().numpy().tobytes()
        return [raw[len(b) * k:len(b) * (k + 1)] for k in range(self.world)]

    def gather_dev(self, vec, nbytes: int):
        from .native import check, lib
        from .spartan import DeviceVec
        out = DeviceVec(nbytes * self.world)
        check(lib().b200_sync())  # the producer kernels of `vec` ran on the library's stream
        src = torch.as_tensor(_CudaView(vec.ptr.value, nbytes), device="cuda")
        dst = torch.as_tensor(_CudaView(out.ptr.value, nbytes * self.world), device="cuda")
        dist.all_gather_into_tensor(dst, src, group=self.group)
        torch.cuda.current_stream().synchronize()
        return out


def sharded_hyperkzg_prove(curve, ck, P_local, x: list, r, q, comm, on_w=None):
    """`EvaluationEngine::prove` of HyperKZG (hyperkzg.rs:926-1116; the single-GPU mirror is
    spartan.hyperkzg_prove_resident) with the polynomial split by INDEX RANGE over comm.world ranks (a power of
    two): rank g holds P[g n/G .. (g+1) n/G) in `P_local` (a DeviceVec).  `ck` is the FULL key, resident on
    every GPU (level i of the fold chain lives on indices [0, n/2^i), so the slices a rank commits move towards
    the front of the key; 3.5 GB of tables at 2^22 against 180 GB of HBM).

      fold        local: the pairs (2j, 2j+1) are adjacent, so a slice of level i folds into the same rank's
                  slice of level i+1 -- until a level has one element per rank, which is all-gathered and the
                  short tail (G elements) continues replicated
      com_i       per-rank partial MSM over ck[lo_i .. hi_i) (b200_msm_dev with a base offset), ONE all-gather of
                  the (levels x 96 B) partials, local sum
      v_i(u_t)    local Horner on the slice, scaled by u_t^lo_i, one all-gather of 3 values per level
      B           = sum_k q^k P_k zero-extended, by index range of P_0: the levels k >= 1 (n - G elements in all)
                  are all-gathered once -- the only bulk exchange -- and every rank combines its own range
      h_t, w_t    division by (X - u_t) on index ranges with a carry from the ranks to the right (one all-gather
                  of 3 values), partial MSMs over ck[lo_0 .. hi_0), one all-gather of 3 partials

    `r`, `q`, `on_w` as in hyperkzg_prove_resident (values or transcript callables; the transcript runs
    replicated on every rank).  Returns (com, v, w) identical on every rank and to the unsharded prover."""
    import ctypes

    from . import fields
    from .native import c_size_t, check, lib
    from .provider import Curve, _jac_to_affine
    from .spartan import DeviceVec
    L = lib()
    curve = Curve(curve)
    fid = curve.scalar_field
    p = fields.MODULUS[fid]
    G, g = comm.world, comm.rank
    ell = len(x)
    n = 1 << ell
    assert G & (G - 1) == 0 and 2 * G <= n, "world must be a power of two with at least two elements per rank"
    nloc = n // G
    base_ptr = lambda v, k: ctypes.c_void_p(v.ptr.value + 32 * k)

    def affine_sum(parts: list):
        """sum of Jacobian points given as 96-byte strings -> affine"""
        d, out = DeviceVec.from_bytes(b"".join(parts)), DeviceVec(96)
        check(L.b200_jacobian_sum_dev(int(curve), d.ptr, len(parts), out.ptr, None))
        raw = out.to_bytes()
        d.free()
        out.free()
        return _jac_to_affine(curve, raw)

    # ---- Phase 1: the fold chain (hyperkzg.rs:1083-1095) -------------------------------------------
    # level i: (vector, local length, sharded?).  Sharded levels hold indices [g len_i/G, (g+1) len_i/G).
    levels = [(P_local, nloc, True)]
    for i in range(ell - 1):
        vec, ln, sharded = levels[-1]
        if sharded and ln == 1:  # one element per rank: replicate the tail
            vec, ln, sharded = comm.gather_dev(vec, 32), G, False
            levels[-1] = (vec, ln, sharded)
        xi = DeviceVec.from_bytes(fields.to_mont_bytes(fid, x[ell - i - 1]))
        nxt = DeviceVec(16 * ln)
        check(L.b200_kzg_fold_dev(fid, vec.ptr, ln, xi.ptr, nxt.ptr, None))
        check(L.b200_sync())  # xi is released on return
        levels.append((nxt, ln // 2, sharded))
    full_len = lambda i: n >> i
    lo_of = lambda i: g * (full_len(i) // G)

    # ---- com_i = commit(P_i), i = 1 .. ell-1 (:1099-1100) ------------------------------------------
    parts = DeviceVec(96 * max(ell - 1, 1))
    if ell > 1:  # all levels in one call: the MSMs are spread over the key's lanes, their latency-bound tails overlap
        k = ell - 1
        offs = (c_size_t * k)(*[lo_of(i) if levels[i][2] else 0 for i in range(1, ell)])  # replicated tail: offset 0
        ptrs = (ctypes.c_void_p * k)(*[levels[i][0].ptr.value for i in range(1, ell)])
        lns = (c_size_t * k)(*[levels[i][1] for i in range(1, ell)])
        check(L.b200_msm_many_dev(ck.handle, offs, ptrs, lns, k, parts.ptr, None))
    mine = parts.to_bytes(96 * (ell - 1))
    allp = comm.gather_bytes(mine) if ell > 1 else [b""]
    com = []
    for i in range(1, ell):
        sl = slice(96 * (i - 1), 96 * i)
        if levels[i][2]:
            com.append(affine_sum([a[sl] for a in allp]))
        else:
            com.append(_jac_to_affine(curve, mine[sl]))
    if callable(r):
        r = r(com)
    u = [r % p, (-r) % p, r * r % p]  # :1105-1106
    us = DeviceVec.from_bytes(fields.pack(fid, u))

    # ---- v[i][t] = P_i(u_t) (:1048-1056) ------------------------------------------------------------
    ev = DeviceVec(96 * ell)
    for i in range(ell):
        vec, ln, _ = levels[i]
        check(L.b200_poly_eval_dev(fid, vec.ptr, ln, us.ptr, 3, ctypes.c_void_p(ev.ptr.value + 96 * i), None))
    loc = [fields.unpack(fid, ev.to_bytes(96 * ell)[96 * i:96 * i + 96]) for i in range(ell)]
    scaled = b"".join(int(loc[i][t] * pow(u[t], lo_of(i), p) % p).to_bytes(32, "little")
                      for i in range(ell) for t in range(3))
    allv = comm.gather_bytes(scaled)
    v = []
    for i in range(ell):
        if levels[i][2]:
            v.append([sum(int.from_bytes(a[32 * (3 * i + t):32 * (3 * i + t) + 32], "little") for a in allv) % p
                      for t in range(3)])
        else:
            v.append(list(loc[i]))
    if callable(q):
        q = q(v)

    # ---- B = sum_k q^k P_k on this rank's index range [lo_0, hi_0) (:1028-1040) ----------------------
    lo0, hi0 = g * nloc, (g + 1) * nloc
    srcs, coef = [base_ptr(P_local, 0)], [1]
    lens = [nloc]
    gathered = []
    for k in range(1, ell):
        vec, ln, sharded = levels[k]
        if full_len(k) <= lo0:
            if sharded:  # still takes part in the collective
                gathered.append(comm.gather_dev(vec, 32 * ln))
            continue
        whole = comm.gather_dev(vec, 32 * ln) if sharded else vec
        if sharded:
            gathered.append(whole)
        srcs.append(base_ptr(whole, lo0))
        lens.append(min(hi0, full_len(k)) - lo0)
        coef.append(pow(q, k, p))
    assert len(srcs) <= 32, "rlc of more than 32 polynomials"
    qd = DeviceVec.from_bytes(fields.pack(fid, coef))
    Bloc = DeviceVec(32 * (nloc + 1))  # one spare slot: the carry coefficient of the division below
    ptrs = (ctypes.c_void_p * len(srcs))(*[s_.value for s_ in srcs])
    lns = (c_size_t * len(srcs))(*lens)
    check(L.b200_rlc_dev(fid, ptrs, lns, len(srcs), qd.ptr, nloc, Bloc.ptr, None))

    # ---- h_t = B / (X - u_t), w_t = commit(h_t) (:1062-1065) ----------------------------------------
    bev = DeviceVec(96)
    check(L.b200_poly_eval_dev(fid, Bloc.ptr, nloc, us.ptr, 3, bev.ptr, None))
    mine_v = fields.unpack(fid, bev.to_bytes(96))
    allb = comm.gather_bytes(b"".join(int(e).to_bytes(32, "little") for e in mine_v))
    wparts = DeviceVec(96 * 3)
    cnt = nloc - 1 if g == G - 1 else nloc  # h has n - 1 coefficients: the last rank owns one less
    keep = []
    for t in range(3):
        carry = 0  # sum over ranks k > g of V_k(u_t) u_t^(lo_k - hi_0)
        for k in range(g + 1, G):
            vk = int.from_bytes(allb[k][32 * t:32 * t + 32], "little")
            carry = (carry + vk * pow(u[t], (k - g - 1) * nloc, p)) % p
        cd = DeviceVec.from_bytes(fields.to_mont_bytes(fid, carry))
        check(L.b200_memcpy_d2d(base_ptr(Bloc, nloc), cd.ptr, 32, None))
        ud = DeviceVec.from_bytes(fields.to_mont_bytes(fid, u[t]))
        h = DeviceVec(32 * nloc)
        check(L.b200_poly_div_dev(fid, Bloc.ptr, nloc + 1, ud.ptr, h.ptr, None))  # -> h[lo_0 .. hi_0)
        keep.append((cd, ud, h))
    offs3 = (c_size_t * 3)(lo0, lo0, lo0)
    ptrs3 = (ctypes.c_void_p * 3)(*[kp[2].ptr.value for kp in keep])
    lns3 = (c_size_t * 3)(cnt, cnt, cnt)
    check(L.b200_msm_many_dev(ck.handle, offs3, ptrs3, lns3, 3, wparts.ptr, None))
    allw = comm.gather_bytes(wparts.to_bytes(96 * 3))  # (the download synchronises: `keep` may go now)
    w = [affine_sum([a[96 * t:96 * t + 96] for a in allw]) for t in range(3)]
    if on_w is not None:
        on_w(w)
    return com, v, w


# ---------------------------------------------------------------------------------------------
# The sharded MSM with its collective fused into the reduction kernel (b200_msm_sharded_dev)
# ---------------------------------------------------------------------------------------------
class PeerGroup:
    """One exchange buffer per rank, mapped into every other rank's process with CUDA IPC (NVLink peer stores):
    the last kernel of each rank's MSM writes its partial sum into all peers' buffers, waits for theirs and adds
    them in a fixed order (csrc/msm_kernels.cuh `peer_exchange_sum`).  torch.distributed is used ONCE, here, to
    exchange the 64-byte IPC handles; the MSM steps themselves make no NCCL call.  All ranks must issue their
    `msm` calls in the same order."""

    def __init__(self, group=None):
        import ctypes

        from .native import c_u64, check, lib
        L = lib()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.buf = ctypes.c_void_p()
        check(L.b200_peer_buffer_alloc(ctypes.byref(self.buf)))
        h = ctypes.create_string_buffer(64)
        check(L.b200_ipc_export(self.buf, h))
        if self.world == 1:
            handles = [h.raw]
        elif dist.get_backend(group) == "nccl":
            handles = NcclComm(group).gather_bytes(h.raw)
        else:
            handles = all_gather_bytes(h.raw, group)
        self.mapped = []
        ptrs = (ctypes.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = self.buf.value
            else:
                p = ctypes.c_void_p()
                check(L.b200_ipc_open(ctypes.create_string_buffer(handles[r], 64), ctypes.byref(p)))
                self.mapped.append(p)
                ptrs[r] = p.value
        g = c_u64(0)
        check(L.b200_peer_group_create(self.rank, self.world, ptrs, ctypes.byref(g)))
        self.handle = g.value
        if dist.is_initialized():
            dist.barrier(group)  # every buffer is mapped everywhere before the first exchange

    def msm(self, ck, base_offset: int, d_scalars: int, n: int, d_out: int, stream=None):
        """sum over ALL ranks of  sum_i scalars_r[i] * ck_r[base_offset + i]  -> d_out (96-byte Jacobian), on every
        rank, bit-identical.  d_scalars / d_out: device addresses; asynchronous on `stream`."""
        import ctypes

        from .native import check, lib
        check(lib().b200_msm_sharded_dev(ck.handle, base_offset, ctypes.c_void_p(d_scalars), n, self.handle,
                                         ctypes.c_void_p(d_out), stream))

    def status(self):
        from .native import check, lib
        check(lib().b200_peer_group_status(self.handle))

    def close(self):
        from .native import check, lib
        L = lib()
        if self.handle:
            check(L.b200_sync())
            if dist.is_initialized():
                dist.barrier()  # nobody unmaps a buffer a peer may still write
            check(L.b200_peer_group_release(self.handle))
            self.handle = 0
            for p in self.mapped:
                L.b200_ipc_close(p)
            L.b200_peer_buffer_free(self.buf)
