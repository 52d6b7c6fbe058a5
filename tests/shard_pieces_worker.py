"""Worker for tests/test_sharding_pieces.py: rank `rank` of `world` over gloo; engine "oracle" (CPU) or "gpu"."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleEngine:
    """CPU stand-in for nova_b200.sharding.DeviceEngine (test infrastructure)."""

    def __init__(self, fid):
        from oracle import coracle
        self.co, self.fid = coracle, fid

    def matrix(self, data, indices, indptr, cols):
        return (data, list(indices), list(indptr))

    def spmv(self, mat, z):
        return self.co.spmv(self.fid, mat[0], mat[1], mat[2], z)

    def vec_add(self, a, b):
        return self.co.vec_add(self.fid, a, b)

    def cross_term(self, az, bz, cz, e1, u, e2=None):
        return self.co.cross_term(self.fid, az, bz, cz, e1, e2, u)

    def poly_eval(self, f, u):
        return self.co.poly_eval(self.fid, f, u)

    def poly_div(self, f, u):
        return self.co.poly_div(self.fid, f, u)


def main():
    rank, world, port, kind, outpath = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nova_b200 import sharding as sh
    from oracle import coracle as co
    from oracle.ppsnark_ref import random_instance
    from oracle.pyref import FIELD_MODULUS, SplitMix64, mont_bytes
    from snark_parity import csr
    fid = 0
    p = FIELD_MODULUS[fid]
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    if kind in ("gpu", "emulated"):
        import nova_b200
        if kind == "emulated":  # the real DeviceEngine adapters, the library answered by the oracle
            import emulated_device
            emulated_device.install()
        from nova_b200.native import check, lib
        check(lib().b200_init(0))
        eng = sh.DeviceEngine(fid)
    else:
        eng = OracleEngine(fid)
    ok = True
    rng = SplitMix64(4711)
    # ---- Horner evaluation and division by (X - u), ragged sizes (n not divisible by world) ----
    for n in (1, 2, 37, 1000, 4097):
        if n < world:
            continue
        f = co.gen_scalars(fid, 100 + n, n)
        u = rng.field(p)
        lo, hi = sh.shard_range(n, rank, world)
        got = sh.sharded_poly_eval(eng, p, f[32 * lo:32 * hi], lo, u)
        exp = int.from_bytes(co.poly_eval(fid, f, mont_bytes(p, u)), "little") * pow(1 << 256, -1, p) % p
        ok &= got == exp
        if n >= 2:
            q_local = sh.sharded_poly_div(eng, p, f[32 * lo:32 * hi], lo, hi, n, u)
            q_all = b"".join(sh.all_gather_var(q_local))
            ok &= q_all == co.poly_div(fid, f, mont_bytes(p, u))
    # ---- folding-step cross term on row slices (commit_T and commit_T_relaxed shapes) ----
    num_cons, num_vars, num_io = 64, 32, 2
    S, W, u1, X1 = random_instance(p, rng, num_cons, num_vars, num_io)
    W2 = [rng.field(p) for _ in range(num_vars)]
    X2 = [rng.field(p) for _ in range(num_io)]
    E2 = [rng.field(p) for _ in range(num_cons)]
    ncols = num_vars + 1 + num_io
    z1, z2 = W["W"] + [u1] + X1, W2 + [1] + X2
    rlo, rhi = sh.shard_range(num_cons, rank, world)   # rows of A, B, C and of E / T
    zlo, zhi = sh.shard_range(ncols, rank, world)      # entries of z
    mats = []
    for name in "ABC":
        rows = [(r - rlo, c, v) for (r, c, v) in S[name] if rlo <= r < rhi]
        d, idx, ptr = csr(rows, rhi - rlo)
        mats.append(eng.matrix(pack(d), idx, ptr, ncols))
    full = [csr(S[name], num_cons) for name in "ABC"]
    Z = pack([(a + b) % p for a, b in zip(z1, z2)])
    az, bz, cz = (co.spmv(fid, pack(d), idx, ptr, Z) for (d, idx, ptr) in full)
    for e2 in (None, E2):
        usum = pack([(u1 + 1) % p])
        T_local = sh.sharded_cross_term(eng, mats, pack(z1[zlo:zhi]), pack(z2[zlo:zhi]), pack(W["E"][rlo:rhi]), usum,
                                        pack(e2[rlo:rhi]) if e2 else None)
        T_exp = co.cross_term(fid, az, bz, cz, pack(W["E"]), pack(e2) if e2 else None, usum)
        ok &= T_local == T_exp[32 * rlo:32 * rhi]
    open(f"{outpath}.{rank}", "w").write("OK" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
