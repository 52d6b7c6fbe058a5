"""Worker for tests/test_ppsnark_sharded.py: rank `rank` of `world` over gloo runs ppsnark.prove_helper_sharded
(the batched inner sum-check of MicroSpartan, ppsnark.rs:886-983) on its CYCLIC shards of the sixteen tables and
compares every prover message with the unsharded oracle restatement.  kind "emulated" or "gpu"."""
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
    if kind == "nccl":  # one GPU per rank, collectives over NVLink
        import torch
        torch.cuda.set_device(rank)
    dist.init_process_group("nccl" if kind == "nccl" else "gloo", rank=rank, world_size=world)
    import nova_b200  # noqa: F401
    if kind == "emulated":
        import emulated_device
        emulated_device.install()
    from nova_b200 import ppsnark as dp
    from nova_b200 import sharding as sh
    from nova_b200 import spartan as sp
    from nova_b200.native import check, lib
    from oracle import ppsnark_ref as pr
    from oracle.pyref import FIELD_MODULUS, Keccak256Transcript, SplitMix64, eq_evals, mont_bytes
    check(lib().b200_init(rank if kind == "nccl" else 0))
    fid = 0
    p = FIELD_MODULUS[fid]
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    ok = True
    for ell, zero_rho, zero_outer in ((4, (), ()), (5, (1,), (0, 3)), (6, (0,), ())):
        N = 1 << ell
        if N < 2 * world:
            continue
        rng = SplitMix64(9000 + ell)
        vec = lambda: [rng.field(p) for _ in range(N)]
        oracles, aux, ts_row, ts_col = [vec() for _ in range(4)], [vec() for _ in range(4)], vec(), vec()
        L_row, L_col, val, E, W = vec(), vec(), vec(), vec(), vec()
        rhos = [0 if i in zero_rho else rng.field(p) for i in range(ell)]
        r_outer = [0 if i in zero_outer else rng.field(p) for i in range(ell)]
        claim, claim_E = rng.field(p), rng.field(p)
        num_vars = 4
        mem = pr.MemorySumcheckInstance(p, oracles, aux, rhos, ts_row, ts_col)
        inner = pr.InnerBatchedSumcheckInstance(p, claim, L_row, L_col, val, claim_E, r_outer, E)
        wit = pr.WitnessBoundSumcheck(p, r_outer, W, num_vars)
        tr_ref = Keccak256Transcript(p, b"shard")
        exp = pr.prove_helper(p, mem, inner, wit, tr_ref)
        # this rank's cyclic shards
        cyc = lambda v: sp.DeviceVec.from_bytes(pack(v[rank::world]))
        nl = N // world
        masked = eq_evals(p, r_outer)
        for i in range(1 << (num_vars.bit_length() - 1)):
            masked[i] = 0  # ppsnark.rs:288-297: the first 2^m entries of eq(tau) are zeroed
        d_mem = dp.MemorySumcheckInstance(fid, nl, [cyc(v) for v in oracles], [cyc(v) for v in aux], rhos, cyc(ts_row),
                                          cyc(ts_col))
        d_inner = dp.InnerBatchedSumcheckInstance(fid, nl, claim, cyc(L_row), cyc(L_col), cyc(val), claim_E, r_outer,
                                                  cyc(E))
        d_wit = dp.WitnessBoundSumcheck.from_shards(fid, nl, cyc(W), cyc(masked))
        tr = Keccak256Transcript(p, b"shard")
        got = dp.prove_helper_sharded(fid, d_mem, d_inner, d_wit, tr, rank, world, sh.all_gather_bytes)
        ok &= [list(q) for q in got[0]] == [list(q) for q in exp[0]] and list(got[1]) == list(exp[1])
        ok &= got[2:] == exp[2:] and tr.squeeze(b"z") == tr_ref.squeeze(b"z")
    dist.barrier()
    open(f"{outpath}.{rank}", "w").write("OK" if ok else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
