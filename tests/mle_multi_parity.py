"""Shared body: MultilinearPolynomial::multi_evaluate_with through nova_b200.spartan.mle_eval_multi_dev, restating
the reference's tests (src/spartan/polys/multilinear.rs:414-484): equal to the single-polynomial evaluation, the
one-polynomial case, and the known values p = (x1 + x2) x3 -> 2 and the constant 5 at (1, 1, 1)."""
from oracle.pyref import FIELD_MODULUS, SplitMix64, mle_evaluate, mont_bytes


def run(sp, fid):
    p = FIELD_MODULUS[fid]
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    up = lambda xs: sp.DeviceVec.from_bytes(pack(xs))
    rng = SplitMix64(4200 + fid)
    for num_vars, k in ((6, 3), (4, 1), (9, 5), (1, 2), (0, 2)):
        n = 1 << num_vars
        polys = [[rng.field(p) for _ in range(n)] for _ in range(k)]
        pt = [rng.field(p) for _ in range(num_vars)]
        r_dev = up(pt) if num_vars else sp.DeviceVec(32)
        got = sp.mle_eval_multi_dev(fid, [up(z) for z in polys], num_vars, r_dev)
        assert got == [mle_evaluate(p, z, pt) for z in polys], (num_vars, k)
    z1, z2 = [0, 0, 0, 1, 0, 1, 0, 2], [5] * 8
    assert sp.mle_eval_multi_dev(fid, [up(z1), up(z2)], 3, up([1, 1, 1])) == [2, 5]
    assert sp.mle_eval_multi_dev(fid, [], 3, up([1, 1, 1])) == []
