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

    def axpy(self, a, b, r):
        return self.co.axpy(self.fid, a, b, r)

    def msm_partial(self, ck, base_offset, scalars):  # ck: (curve id, bases); affine result as a Jacobian z = 1
        from oracle.pyref import CURVES, mont_bytes
        cid, bases = ck
        n = len(scalars) // 32
        aff = self.co.msm(cid, scalars, bases[64 * base_offset:64 * (base_offset + n)]) if n else bytes(64)
        return aff + (bytes(32) if aff == bytes(64) else mont_bytes(CURVES[cid].p, 1))

    def point_sum(self, parts):
        from oracle.pyref import CURVES
        c = CURVES[0]
        acc = None
        for b in parts:
            acc = c.add(acc, c.jacobian_from_bytes(b))
        return acc


def main():
    rank, world, port, kind, outpath = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if kind == "nccl":  # one GPU per rank, collectives over NVLink
        import torch
        torch.cuda.set_device(rank)
    dist.init_process_group("nccl" if kind == "nccl" else "gloo", rank=rank, world_size=world)
    from nova_b200 import sharding as sh
    from oracle import coracle as co
    from oracle.ppsnark_ref import random_instance
    from oracle.pyref import FIELD_MODULUS, SplitMix64, mont_bytes
    from snark_parity import csr
    fid = 0
    p = FIELD_MODULUS[fid]
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    if kind in ("gpu", "emulated", "nccl"):
        import nova_b200
        if kind == "emulated":  # the real DeviceEngine adapters, the library answered by the oracle
            import emulated_device
            emulated_device.install()
        from nova_b200.native import check, lib
        check(lib().b200_init(rank if kind == "nccl" else 0))
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
    # ---- the folding half of prove_step over the ranks (commit W2, commit_T, folds) ----
    from oracle.pyref import CURVES
    cid, c = 0, CURVES[0]
    n_key = max(num_cons, num_vars)
    bases = co.gen_bases(cid, n_key)
    if kind in ("gpu", "emulated", "nccl"):
        from nova_b200 import provider
        eng.curve = cid
        ck = provider.CommitmentKey(provider.Curve(cid), bases)
    else:
        ck = (cid, bases)
    wlo, whi = sh.shard_range(num_vars, rank, world)
    seen = []

    def challenge(comm_W2, comm_T):
        seen.append((comm_W2, comm_T))
        return 0x1234567 * (1 + (comm_T[0] % 97))  # any function of the messages: must agree on every rank
    usum = pack([(u1 + 1) % p])
    comm_W2, comm_T, r, W_loc, E_loc, T_loc = sh.sharded_fold_step(
        eng, p, ck, mats, (rlo, rhi), (wlo, whi), pack(z1[zlo:zhi]), pack(z2[zlo:zhi]), pack(W["W"][wlo:whi]),
        pack(W["E"][rlo:rhi]), pack(W2[wlo:whi]), usum, challenge)
    T_exp = co.cross_term(fid, az, bz, cz, pack(W["E"]), None, usum)
    aff = c.affine_from_bytes
    ok &= comm_W2 == aff(co.msm(cid, pack(W2), bases)) and comm_T == aff(co.msm(cid, T_exp, bases))
    ok &= T_loc == T_exp[32 * rlo:32 * rhi]
    rm = mont_bytes(p, r)
    ok &= W_loc == co.axpy(fid, pack(W["W"]), pack(W2), rm)[32 * wlo:32 * whi]
    ok &= E_loc == co.axpy(fid, pack(W["E"]), T_exp, rm)[32 * rlo:32 * rhi]
    ok &= len(seen) == 1
    open(f"{outpath}.{rank}", "w").write("OK" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
