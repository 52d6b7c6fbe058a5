"""Worker for tests/test_ptau_sharded.py: rank `rank` of `world` over gloo loads ITS slice of a PTAU file
(nova_b200.ptau.load_setup_sharded), commits through the slice keys and checks the error agreement.
kind "emulated": the library answered by the oracle (CPU); "gpu": the real library (every rank on device 0)."""
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, port, kind, path, outpath = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4],
                                              sys.argv[5], sys.argv[6])
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nova_b200  # noqa: F401
    if kind == "emulated":
        import emulated_device
        emulated_device.install()
    from nova_b200 import ptau
    from nova_b200.native import check, lib
    from oracle import coracle as co
    from oracle.pyref import CURVES, SplitMix64, mont_bytes
    check(lib().b200_init(0))
    cid, c = 0, CURVES[0]
    p = c.q
    n = 100  # -> 128 points
    h = co.gen_bases(cid, 1, 777)
    raw = open(path, "rb").read()
    srs = raw[raw.index(co.scalar_mul(cid, c.affine_bytes(c.gen), 1)):][:64 * 128]
    ok = True
    with open(path, "rb") as f:
        ck, lo, hi = ptau.load_setup_sharded(f, h, n, rank, world)
    ok &= (lo, hi) == ((128 * rank) // world, (128 * (rank + 1)) // world) and ck.n == hi - lo
    ok &= ck.bases == srs[64 * lo:64 * hi] and (ck.h is not None) == (rank == 0)
    rng = SplitMix64(5)
    v = [rng.field(p) for _ in range(n)]
    r = rng.field(p)
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    got = ptau.sharded_commit(ck, lo, hi, pack(v[lo:min(hi, n)]), mont_bytes(p, r), rank)
    exp = c.add(c.msm_naive(v, [c.affine_from_bytes(srs[64 * i:64 * i + 64]) for i in range(n)]),
                c.mul(r, c.affine_from_bytes(h)))
    ok &= got == exp
    got0 = ptau.sharded_commit(ck, lo, hi, pack(v[lo:min(hi, n)]), None, rank)
    ok &= got0 == c.msm_naive(v, [c.affine_from_bytes(srs[64 * i:64 * i + 64]) for i in range(n)])
    ck.release()
    # error agreement: two corrupted points in different ranks' ranges -> every rank names the smaller index
    g1_off = raw.index(srs[:64])
    for bad_pts, exc, needle in (((101, 17), ptau.PointNotOnCurve, "17"), ((120,), ptau.PointNotOnCurve, "120")):
        broken = bytearray(raw)
        for i in bad_pts:
            broken[g1_off + 64 * i + 33] ^= 4
        try:
            ptau.load_setup_sharded(io.BytesIO(bytes(broken)), h, n, rank, world)
            ok = False
        except exc as e:
            ok &= needle in str(e)
    # a non-canonical coordinate in the last rank's range: io error on every rank
    broken = bytearray(raw)
    x = int.from_bytes(broken[g1_off + 64 * 127:g1_off + 64 * 127 + 32], "little") + c.p
    broken[g1_off + 64 * 127:g1_off + 64 * 127 + 32] = x.to_bytes(32, "little")
    try:
        ptau.load_setup_sharded(io.BytesIO(bytes(broken)), h, n, rank, world)
        ok = False
    except ptau.IoError as e:
        ok &= "127" in str(e)
    # an invalid blinding generator (held by rank 0 only) stops every rank
    try:
        ptau.load_setup_sharded(io.BytesIO(raw), h[:32] + bytes(32), n, rank, world)
        ok = False
    except ptau.PointNotOnCurve as e:
        ok &= "blinding" in str(e)
    dist.barrier()
    open(f"{outpath}.{rank}", "w").write("OK" if ok else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
