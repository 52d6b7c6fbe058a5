"""Worker for the sharded sum-check tests: rank `rank` of `world` (gloo rendezvous on 127.0.0.1).
engine = "oracle" (CPU, any box) or "gpu" (every rank uses cuda:0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleEngine:
    """CPU stand-in for the per-rank device engine (test infrastructure)."""

    def __init__(self, fid):
        from oracle import coracle, pyref
        self.co, self.pyref, self.fid = coracle, pyref, fid
        self.p = pyref.FIELD_MODULUS[fid]

    def upload(self, b):
        return bytearray(b)

    def download(self, h, n):
        return bytes(h[:32 * n])

    def download_canonical(self, h):
        return self.pyref.from_mont_bytes(self.p, bytes(h[:32])).to_bytes(32, "little")

    def eq_tables(self, taus):
        inst = self.pyref.EqSumCheckInstance(self.p, taus)
        pk = lambda xs: b"".join(self.pyref.mont_bytes(self.p, x) for x in xs)

        class _T:
            def tables(_, rnd):
                inst.round = rnd
                L, R, sh = inst.tables()
                return (pk(L) if L else None), pk(R), sh
        return _T()

    def sc_eval(self, form, A, B, C, local_len, left, right, shift, id_mul, id_add):
        raw = self.co.sc_eval(self.fid, form, bytes(A[:32 * local_len]), bytes(B[:32 * local_len]),
                              bytes(C[:32 * local_len]), left, right, shift, id_mul, id_add)
        return [self.pyref.from_mont_bytes(self.p, raw[i:i + 32]) for i in range(0, len(raw), 32)]

    def bind(self, h, local_len, r):
        out = self.co.bind_top(self.fid, bytes(h[:32 * local_len]), self.pyref.mont_bytes(self.p, r))
        h[:len(out)] = out


def main():
    rank, world, port, engine_kind, fid, l, zero_tau, outpath = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]),
                                                                  sys.argv[4], int(sys.argv[5]), int(sys.argv[6]),
                                                                  int(sys.argv[7]), sys.argv[8])
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if engine_kind == "nccl":  # one GPU per rank
        import torch
        torch.cuda.set_device(rank)
    dist.init_process_group("nccl" if engine_kind == "nccl" else "gloo", rank=rank, world_size=world)
    from nova_b200.sharding import cyclic_shard, sharded_prove_cubic_with_three_inputs
    from oracle.pyref import FIELD_MODULUS, Keccak256Transcript, SplitMix64, mont_bytes, prove_cubic_with_three_inputs
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(1000 + l)
    n = 1 << l
    A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
    taus = [rng.field(p) for _ in range(l)]
    if zero_tau:
        taus[0] = 0
        taus[l - 1] = 0
    claim = rng.field(p)
    if engine_kind in ("gpu", "nccl"):
        import nova_b200  # noqa: F401
        from nova_b200.native import check, lib
        from nova_b200.spartan import DeviceSumcheckEngine
        check(lib().b200_init(rank if engine_kind == "nccl" else 0))
        eng = DeviceSumcheckEngine(fid)
    else:
        eng = OracleEngine(fid)
    pk = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    hs = [eng.upload(cyclic_shard(pk(v), rank, world)) for v in (A, B, C)]
    got = sharded_prove_cubic_with_three_inputs(eng, p, claim, taus, hs[0], hs[1], hs[2],
                                                Keccak256Transcript(p, b"sc"), rank, world)
    exp = prove_cubic_with_three_inputs(p, claim, taus, A, B, C, Keccak256Transcript(p, b"sc"))
    ok = got == exp
    with open(f"{outpath}.{rank}", "w") as f:
        f.write("OK" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
