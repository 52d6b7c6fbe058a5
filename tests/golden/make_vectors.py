#!/usr/bin/env python3
"""Generates tests/golden/oracle_vectors.json -- small seeded input/output vectors for the hot path.

The reference (Rust) cannot be built or imported in this image, so the outputs here come from the
PYTHON BIG-INTEGER oracle (oracle/pyref.py: plain integers mod p, affine group law, Keccak
transcript), which is itself pinned by the reference's literal vectors in reference_kats.json.
Nothing in this script touches the C restatement or the CUDA library: both are CHECKED against the
file it writes (tests/test_golden_fixtures.py), the C oracle on CPU and the device through the C ABI.

Inputs are derived from SplitMix64 seeds (our generator, documented in oracle/pyref.py), so the file
stores seeds + expected outputs, not megabytes of inputs.   Run:  python tests/golden/make_vectors.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import pyref  # noqa: E402
from oracle.pyref import CURVES, FIELD_MODULUS, Keccak256Transcript, SplitMix64  # noqa: E402


def hx(x: int) -> str:
    return pyref.to_repr(x).hex()


def pt(P):
    return None if P is None else [hx(P[0]), hx(P[1])]


def commit_cases():
    """commit(ck, v, r) = MSM(v, ck[..n]) + r*h  (pedersen.rs:263-270, hyperkzg.rs:584-591) with
    ck = bases_arith(n+1): P_i = (k0 + i) * G, h = the last point; scalars: uniform, then the
    reference's structured shapes (all-equal, alternating 0/(q-1): curve_property_tests.rs:196-215)."""
    out = []
    for cid in range(4):
        c = CURVES[cid]
        n = 37
        bases = c.bases_arith(n + 1)
        rng = SplitMix64(1000 + cid)
        uni = [rng.field(c.q) for _ in range(n)]
        eq = [uni[0]] * n
        alt = [0 if i % 2 == 0 else c.q - 1 for i in range(n)]
        small = [(i * 37 + 1) % 1024 for i in range(n)]
        r = rng.field(c.q)
        for kind, v, blind in (("uniform", uni, r), ("uniform_r0", uni, 0), ("all_equal", eq, 0), ("alternating", alt, 0),
                               ("u10", small, 0), ("empty", [], 0)):
            acc = c.msm_naive(v, bases[:len(v)])
            if blind:
                acc = c.add(acc, c.mul(blind, bases[n]))
            out.append({"curve": cid, "kind": kind, "n": len(v), "key_len": n, "seed": 1000 + cid,
                        "scalars": [hx(s) for s in v], "blind": hx(blind), "commitment": pt(acc)})
    return out


def field_vector_cases():
    """T = Az o Bz - u*Cz - E (r1cs/mod.rs:614-620), W1 + r*W2 (r1cs/mod.rs:1044-1064), bind_poly_var_top
    (multilinear.rs:65-84) on n = 10 (ragged for the vectorised kernels)."""
    out = []
    for fid in range(4):
        p = FIELD_MODULUS[fid]
        rng = SplitMix64(2000 + fid)
        n = 10
        az, bz, cz, e = ([rng.field(p) for _ in range(n)] for _ in range(4))
        u, r = rng.field(p), rng.field(p)
        T = [(a * b - u * cc - ee) % p for a, b, cc, ee in zip(az, bz, cz, e)]
        fold = [(a + r * b) % p for a, b in zip(az, bz)]
        z8 = az[:8]
        out.append({"field": fid, "seed": 2000 + fid, "n": n,
                    "az": [hx(x) for x in az], "bz": [hx(x) for x in bz], "cz": [hx(x) for x in cz],
                    "e": [hx(x) for x in e], "u": hx(u), "r": hx(r),
                    "cross_term": [hx(x) for x in T], "fold": [hx(x) for x in fold],
                    "bind_top_of_az8": [hx(x) for x in pyref.bind_top(p, z8, r)]})
    return out


def sumcheck_cases():
    """Whole sum-check proofs through the Keccak transcript (sumcheck.rs:199-242, 446-507): every
    compressed round polynomial, challenge and final evaluation."""
    out = []
    for fid, l in ((0, 5), (3, 4)):
        p = FIELD_MODULUS[fid]
        rng = SplitMix64(3000 + fid)
        n = 1 << l
        A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
        taus = [rng.field(p) for _ in range(l)]
        claim_q = sum(a * b for a, b in zip(A, B)) % p
        tr = Keccak256Transcript(p, b"golden")
        polys, rs, finals = pyref.prove_quad_prod(p, claim_q, l, A, B, tr)
        after = tr.squeeze(b"end")
        out.append({"kind": "quad_prod", "field": fid, "seed": 3000 + fid, "num_rounds": l, "claim": hx(claim_q),
                    "transcript_label": "golden", "polys": [[hx(x) for x in q] for q in polys],
                    "challenges": [hx(x) for x in rs], "finals": [hx(x) for x in finals],
                    "squeeze_after": hx(after)})
        eq = pyref.eq_evals(p, taus)
        claim_c = sum(e * (a * b - cc) for e, a, b, cc in zip(eq, A, B, C)) % p
        tr = Keccak256Transcript(p, b"golden")
        polys, rs, finals = pyref.prove_cubic_with_three_inputs(p, claim_c, taus, A, B, C, tr)
        after = tr.squeeze(b"end")
        out.append({"kind": "cubic_with_three_inputs", "field": fid, "seed": 3000 + fid, "num_rounds": l,
                    "claim": hx(claim_c), "taus": [hx(x) for x in taus], "transcript_label": "golden",
                    "polys": [[hx(x) for x in q] for q in polys], "challenges": [hx(x) for x in rs],
                    "finals": [hx(x) for x in finals], "squeeze_after": hx(after)})
    return out


def main():
    doc = {
        "_comment": "Generated by tests/golden/make_vectors.py from the Python big-integer oracle (oracle/pyref.py). "
                    "Field elements: hex of the 32-byte little-endian canonical encoding (to_repr). "
                    "Points: affine [x, y] or null for the identity.",
        "generator": {"prng": "SplitMix64 (oracle/pyref.py), field(p) = 64 bytes LE mod p", "bases": "P_i = (0x5EED + i) * G"},
        "commit": commit_cases(),
        "field_vectors": field_vector_cases(),
        "sumcheck": sumcheck_cases(),
    }
    path = os.path.join(HERE, "oracle_vectors.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
