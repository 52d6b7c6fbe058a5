"""Worker for tests/test_hyperkzg_sharded.py: rank `rank` of `world` over gloo runs
nova_b200.sharding.sharded_hyperkzg_prove on its index range of the polynomial and compares every prover
message with the unsharded oracle restatement.  kind "emulated" (library answered by the oracle), "gpu" (gloo, every
rank on device 0) or "nccl" (one GPU per rank, device-to-device collectives)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, port, kind, outpath = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    backend = "nccl" if kind == "nccl" else "gloo"
    if kind == "nccl":  # one GPU per rank
        import torch
        torch.cuda.set_device(rank)
    dist.init_process_group(backend, rank=rank, world_size=world)
    import nova_b200
    if kind == "emulated":
        import emulated_device
        emulated_device.install()
    from nova_b200 import sharding as sh
    from nova_b200 import spartan as sp
    from nova_b200.native import check, lib
    from oracle import coracle as co
    from oracle import hyperkzg_ref as hk
    from oracle.pyref import CURVES, Keccak256Transcript, SplitMix64
    check(lib().b200_init(rank if kind == "nccl" else 0))
    cid, c = 0, CURVES[0]
    fid, p = c.scalar_field, c.q
    comm = sh.NcclComm() if kind == "nccl" else sh.HostStagedComm()
    ok = True
    # ell chosen so that: the tail is replicated for several levels (ell = 6, world 4: levels of 4, 2 elements),
    # the last level still has one element per rank (ell = 2 with world 2), and a larger ragged-free case
    for ell in (2, 3, 6, 9):
        n = 1 << ell
        if 2 * world > n:
            continue
        bases = co.gen_bases(cid, n)
        ck = nova_b200.CommitmentKey(nova_b200.Curve(cid), bases)
        hat_P = co.gen_scalars(fid, 40 + ell, n)
        rng = SplitMix64(400 + ell)
        x = [rng.field(p) for _ in range(ell)]
        r, q = rng.field(p), rng.field(p)
        lo, hi = rank * (n // world), (rank + 1) * (n // world)
        P_local = sp.DeviceVec.from_bytes(hat_P[32 * lo:32 * hi])
        got = sh.sharded_hyperkzg_prove(cid, ck, P_local, x, r, q, comm)
        ok &= got == hk.prove_core(cid, bases, hat_P, x, r, q)
        # transcript-driven: challenges derived on every rank from the gathered messages
        tr, tr_ref = Keccak256Transcript(p, b"shard"), Keccak256Transcript(p, b"shard")

        def r_of(com):
            tr.absorb_bytes(b"c", b"".join(sp._commitment_bytes(C) for C in com))
            return tr.squeeze(b"c")

        def q_of(v):
            tr.absorb_bytes(b"v", b"".join(int(e).to_bytes(32, "little") for row in v for e in row))
            return tr.squeeze(b"r")

        def after_w(w):
            tr.absorb_bytes(b"W", b"".join(sp._commitment_bytes(C) for C in w))
            tr.squeeze(b"d")
        P_local = sp.DeviceVec.from_bytes(hat_P[32 * lo:32 * hi])
        com, v, w = sh.sharded_hyperkzg_prove(cid, ck, P_local, x, r_of, q_of, comm, after_w)
        com_r, w_r, v_r = hk.prove(cid, bases, hat_P, x, tr_ref)
        ok &= (com, v, w) == (com_r, v_r, w_r) and tr.squeeze(b"z") == tr_ref.squeeze(b"z")
        ck.release()
    dist.barrier()
    open(f"{outpath}.{rank}", "w").write("OK" if ok else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
