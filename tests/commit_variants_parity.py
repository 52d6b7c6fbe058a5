"""Shared body: the remaining CommitmentEngine methods of the provider mirror (pedersen.rs:285-305, 396-427)
-- commit_small_range, commit_sparse_binary and commit_sparse, each with and without the blind -- against the
oracle.  GPU: tests/test_zz_new_paths_gpu.py; CPU (emulated device): tests/test_provider_mirror_cpu.py."""
from oracle.pyref import CURVES, SplitMix64, mont_bytes


def run(nb, oracle, cid):
    c = CURVES[cid]
    p = c.q
    n = 300
    bases = oracle.gen_bases(cid, n + 1)
    ck = nb.CommitmentKey(nb.Curve(cid), bases[:64 * n], bases[64 * n:])
    eng = nb.CommitmentEngine(cid)
    rng = SplitMix64(31 + cid)
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    r = rng.field(p)
    aff = c.affine_from_bytes

    def ref(scalars, idx, blind):
        sc = pack(list(scalars) + ([blind] if blind else []))
        bs = b"".join(bases[64 * i:64 * i + 64] for i in idx) + (bases[64 * n:] if blind else b"")
        return aff(oracle.msm(cid, sc, bs))

    v = [rng.next() & 0x3FF for _ in range(n)]  # 10-bit integers
    for blind in (0, r):
        rb = pack([blind]) if blind else None
        lo, hi = 37, 250
        assert eng.commit_small_range(ck, v, rb, lo, hi, 10) == ref(v[lo:hi], range(lo, hi), blind)
        ones = sorted({rng.next() % n for _ in range(60)})
        assert eng.commit_sparse_binary(ck, ones, rb) == ref([1] * len(ones), ones, blind)
        idx = [rng.next() % n for _ in range(45)]  # repeated indices are allowed
        sc = [rng.field(p) for _ in idx]
        assert eng.commit_sparse(ck, idx, pack(sc), rb) == ref(sc, idx, blind)
    assert eng.commit_sparse(ck, [], b"", None) is None  # empty -> identity
    # commit_small with a blind, batch_commit_small over ragged vectors (traits.rs:105-116)
    assert eng.commit_small(ck, v, 8, pack([r])) == ref(v, range(n), r)
    vs = [v[:k] for k in (0, 1, 17, n)]
    assert eng.batch_commit_small(ck, vs) == [ref(x, range(len(x)), 0) if x else None for x in vs]
    ck.release()
