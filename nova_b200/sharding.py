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
