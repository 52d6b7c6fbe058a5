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
    t = torch.frombuffer(bytearray(b), dtype=torch.uint8)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [bytes(x.numpy().tobytes()) for x in out]


def sharded_prove_cubic_with_three_inputs(engine, p: int, claim: int, taus: list, A, B, C, transcript,
                                          rank: int, world: int, group=None):
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
    """
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
            parts = [all_gather_bytes(engine.download(h, 1), group) for h in (A, B, C)]
            A, B, C = (engine.upload(b"".join(pp)) for pp in parts)
            local_len, id_mul, id_add = world, 1, 0
        left, right, shift = eq.tables(rnd)

        def gsum(form):
            vals = engine.sc_eval(form, A, B, C, local_len, left, right, shift, id_mul, id_add)
            if id_mul == 1:
                return vals
            raw = b"".join(int(v).to_bytes(32, "little") for v in vals)
            allv = all_gather_bytes(raw, group)
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
