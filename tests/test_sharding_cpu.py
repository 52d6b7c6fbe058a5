"""world_size-2 gloo test of the N>1 host logic (index-range sharding + all-gather of the
per-rank partial points + local combine).  The per-rank MSM is played by the oracle here (no GPU
in this container); on the GPU box bench.py --gpus N runs the same exchange over NCCL."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    from nova_b200.sharding import all_gather_partials, shard_range
    from oracle import coracle as co
    from oracle.pyref import CURVES, mont_bytes

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cid = 0
    c = CURVES[cid]
    lo, hi = shard_range(n, rank, world)
    # every rank owns bases[lo:hi] (generated directly for its slice) and gets scalars[lo:hi]
    bases = co.gen_bases(cid, hi - lo, k0=co.K0_DEFAULT + lo)
    scalars = co.gen_scalars(c.scalar_field, 7, n)[32 * lo:32 * hi]
    part_aff = c.affine_from_bytes(co.msm(cid, scalars, bases, 2))
    p = c.p
    if part_aff is None:
        jac = bytes(96)
    else:  # a non-trivial Jacobian representative (z != 1) to exercise normalisation
        z = 5 + rank
        jac = mont_bytes(p, part_aff[0] * z * z) + mont_bytes(p, part_aff[1] * z ** 3) + mont_bytes(p, z)
    t = torch.frombuffer(bytearray(jac), dtype=torch.uint8)
    allp = all_gather_partials(t)
    total = None
    raw = bytes(allp.numpy().tobytes())
    for r in range(world):
        total = c.add(total, c.jacobian_from_bytes(raw[96 * r:96 * r + 96]))
    if rank == 0:
        full_b = co.gen_bases(cid, n)
        full_s = co.gen_scalars(c.scalar_field, 7, n)
        q.put(total == c.affine_from_bytes(co.msm(cid, full_s, full_b, 2)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1000), (2, 1001), (3, 50)])
def test_sharded_msm_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_tile():
    from nova_b200.sharding import shard_range
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 3):
        for w in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
