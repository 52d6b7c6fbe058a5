"""Shared body: SumcheckProof::prove_batched_cubic (sumcheck.rs:513-577) through the mirror against the oracle
restatement (GPU: tests/test_zz_new_paths_gpu.py; CPU, emulated device: tests/test_spartan_mirror_cpu.py)."""
from oracle.pyref import FIELD_MODULUS, Keccak256Transcript, SplitMix64, mont_bytes, prove_batched_cubic


def run(sp, fid, k, l, zero_tau_at=()):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(8800 + 13 * k + l + fid)
    n = 1 << l
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    As, Bs, Cs = ([[rng.field(p) for _ in range(n)] for _ in range(k)] for _ in range(3))
    alphas = [rng.field(p) for _ in range(k)]
    taus = [0 if i in zero_tau_at else rng.field(p) for i in range(l)]
    claim = rng.field(p)  # the prover is deterministic for any claim
    exp = prove_batched_cubic(p, claim, taus, As, Bs, Cs, alphas, Keccak256Transcript(p, b"bc"))
    got = sp.SumcheckProof.prove_batched_cubic(fid, claim, taus, [pack(v) for v in As], [pack(v) for v in Bs],
                                               [pack(v) for v in Cs], alphas, Keccak256Transcript(p, b"bc"))
    assert [list(q) for q in got[0]] == [list(q) for q in exp[0]]
    assert list(got[1]) == list(exp[1])
    assert [list(c) for c in got[2]] == [list(c) for c in exp[2]]
