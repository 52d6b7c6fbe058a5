"""Worker for the fused sharded MSM (b200_msm_sharded_dev / nova_b200.sharding.PeerGroup): rank `rank` of `world`
commits its index range and must end with the closed-form result of the WHOLE vector, identical on every rank, for
several consecutive epochs (the two slot sets alternate) and for a rank with an empty range.
kind "gpu": gloo, every rank on cuda:0 (CUDA IPC between processes on one device); "nccl": one GPU per rank."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, port, kind, outpath = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if kind == "nccl" else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if kind == "nccl" else "gloo", rank=rank, world_size=world)
    import nova_b200 as nb
    from nova_b200 import sharding as sh
    from nova_b200.native import check, lib
    from nova_b200.provider import Curve, _jac_to_affine
    from oracle import coracle as co
    from oracle.pyref import CURVES
    L = lib()
    check(L.b200_init(dev))
    pg = sh.PeerGroup()
    ok = True
    K0 = 0x5EED
    for cid, n_total in ((0, 1 << 14), (0, 3000), (1, 1 << 12), (0, world - 1)):  # the last: one rank has no pairs
        c = CURVES[cid]
        lo, hi = sh.shard_range(n_total, rank, world)
        n = hi - lo
        sc_all = co.gen_scalars(c.scalar_field, 90 + n_total % 7, n_total)
        ck = nb.CommitmentKey.setup_synthetic(nb.Curve(cid), max(n, 1), k0=K0 + lo)
        d_sc = torch.frombuffer(bytearray(sc_all[32 * lo:32 * hi] or bytes(32)), dtype=torch.uint8).cuda()
        d_out = torch.zeros(96, dtype=torch.uint8, device="cuda")
        k = co.dot_index(c.scalar_field, sc_all, K0)
        exp = c.affine_from_bytes(co.scalar_mul(cid, c.affine_bytes(c.gen), k)) if n_total else None
        for _ in range(3):  # three epochs on the same group
            d_out.zero_()
            pg.msm(ck, 0, d_sc.data_ptr(), n, d_out.data_ptr())
            check(L.b200_sync())
            raw = bytes(d_out.cpu().numpy().tobytes())
            ok &= _jac_to_affine(Curve(cid), raw) == exp
            # every rank must hold the same coordinates, not only the same point
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            t = t.cuda() if kind == "nccl" else t
            allr = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            ok &= all(bytes(x.cpu().numpy().tobytes()) == raw for x in allr)
        ck.release()
    pg.status()
    pg.close()
    open(f"{outpath}.{rank}", "w").write("OK" if ok else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
