"""Committed golden fixtures (tests/golden/):

* reference_kats.json -- literals transcribed from the reference's own tests (file:line inside);
  they pin the Python big-integer oracle and the C restatement               [CPU]
* oracle_vectors.json -- seeded input/output vectors written by tests/golden/make_vectors.py from
  the Python big-integer oracle; the C restatement must reproduce them       [CPU]
  and so must the CUDA path through the C ABI                                [GPU]
"""
import json
import os

import pytest

from oracle.pyref import (CURVES, FIELD_MODULUS, Keccak256Transcript, SplitMix64, from_mont_bytes, keccak256,
                          mont_bytes, to_repr)

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


KATS = _load("reference_kats.json")
VEC = _load("oracle_vectors.json")


def unhex(h):
    return int.from_bytes(bytes.fromhex(h), "little")


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


def ints(p, b):
    return [from_mont_bytes(p, b[i:i + 32]) for i in range(0, len(b), 32)]


# ------------------------------------------------------------------ reference literals (CPU) ---
def test_reference_keccak_vectors():
    k = KATS["keccak_example"]
    assert keccak256(bytes.fromhex(k["input_hex"])).hex() == k["digest_hex"]
    for case in KATS["keccak_transcript"]["cases"]:
        p = int(case["scalar_modulus_hex"], 16)
        t = Keccak256Transcript(p, b"test")
        t.absorb_scalar(b"s1", 2)
        t.absorb_scalar(b"s2", 5)
        assert to_repr(t.squeeze(b"c1")).hex() == case["c1"], case["engine"]
        t.absorb_scalar(b"s3", 128)
        assert to_repr(t.squeeze(b"c2")).hex() == case["c2"], case["engine"]


def test_reference_small_kats(oracle, pyref):
    fid = 0
    p = FIELD_MODULUS[fid]
    s = KATS["spmv"]
    out = oracle.spmv(fid, oracle.field_from_u64(fid, s["csr"]["data"]), s["csr"]["indices"], s["csr"]["indptr"],
                      oracle.field_from_u64(fid, s["z"]))
    assert ints(p, out) == s["expected"]
    dense = [sum(a * z for a, z in zip(row, s["z"])) for row in s["dense"]]
    assert dense == s["expected"]
    for c in KATS["mle"]["cases"]:
        assert pyref.mle_evaluate(p, c["table"], c["point"]) == c["eval"]
        assert ints(p, oracle.mle_eval(fid, oracle.field_from_u64(fid, c["table"]),
                                       oracle.field_from_u64(fid, c["point"]))) == [c["eval"]]
    e = KATS["eq_table"]
    hot = [1 if i == e["one_hot_index"] else 0 for i in range(1 << len(e["r"]))]
    assert pyref.eq_evals(p, e["r"]) == hot
    assert ints(p, oracle.eq_table(fid, oracle.field_from_u64(fid, e["r"]))) == hot
    h = KATS["hyperkzg_eval"]
    for c in h["accept"]:
        assert pyref.mle_evaluate(p, h["poly"], c["point"]) == c["eval"]
    for c in h["reject"]:
        assert pyref.mle_evaluate(p, h["poly"], c["point"]) != c["eval"]


# ------------------------------------------------------------------ engines --------------------
class OracleEngine:
    """The C restatement (oracle/oracle.c) behind the same small interface as the device engine."""

    def __init__(self, oracle):
        self.o = oracle

    def commit(self, cid, bases, key_len, v, blind):
        c = CURVES[cid]
        bb = b"".join(c.affine_bytes(P) for P in bases)
        sc, bs = pack(c.q, v), bb[:64 * len(v)]
        if blind:
            sc, bs = sc + mont_bytes(c.q, blind), bs + bb[64 * key_len:64 * key_len + 64]
        return c.affine_from_bytes(self.o.msm(cid, sc, bs))

    def cross_term(self, fid, az, bz, cz, e, u):
        return self.o.cross_term(fid, az, bz, cz, e, None, u)

    def fold(self, fid, a, b, r):
        return self.o.axpy(fid, a, b, r)

    def bind_top(self, fid, z, r):
        return self.o.bind_top(fid, z, r)

    def prove_quad_prod(self, fid, claim, l, A, B, tr):
        from oracle import pyref
        return pyref.prove_quad_prod(FIELD_MODULUS[fid], claim, l, A, B, tr)

    def prove_cubic3(self, fid, claim, taus, A, B, C, tr):
        from oracle import pyref
        return pyref.prove_cubic_with_three_inputs(FIELD_MODULUS[fid], claim, taus, A, B, C, tr)


class DeviceEngine:
    """The CUDA path through the C ABI (nova_b200 host mirror)."""

    def __init__(self, nb):
        self.nb = nb

    def commit(self, cid, bases, key_len, v, blind):
        c = CURVES[cid]
        bb = b"".join(c.affine_bytes(P) for P in bases)
        ck = self.nb.CommitmentKey(self.nb.Curve(cid), bb[:64 * key_len], bb[64 * key_len:64 * key_len + 64])
        try:
            return self.nb.CommitmentEngine(cid).commit(ck, pack(c.q, v), mont_bytes(c.q, blind) if blind else None)
        finally:
            ck.release()

    def cross_term(self, fid, az, bz, cz, e, u):
        return self.nb.cross_term(fid, az, bz, cz, e, u)

    def fold(self, fid, a, b, r):
        return self.nb.fold_witness(fid, a, b, r)

    def bind_top(self, fid, z, r):
        return self.nb.bind_poly_var_top(fid, z, r)

    def prove_quad_prod(self, fid, claim, l, A, B, tr):
        from nova_b200 import spartan
        p = FIELD_MODULUS[fid]
        return spartan.SumcheckProof.prove_quad_prod(fid, claim, l, pack(p, A), pack(p, B), tr)

    def prove_cubic3(self, fid, claim, taus, A, B, C, tr):
        from nova_b200 import spartan
        p = FIELD_MODULUS[fid]
        return spartan.SumcheckProof.prove_cubic_with_three_inputs(fid, claim, taus, pack(p, A), pack(p, B),
                                                                   pack(p, C), tr)


def _check_commits(eng):
    for case in VEC["commit"]:
        cid, n, key_len = case["curve"], case["n"], case["key_len"]
        c = CURVES[cid]
        bases = c.bases_arith(key_len + 1)
        v = [unhex(h) for h in case["scalars"]]
        exp = None if case["commitment"] is None else tuple(unhex(h) for h in case["commitment"])
        got = eng.commit(cid, bases, key_len, v, unhex(case["blind"]))
        assert (None if got is None else tuple(got)) == exp, (cid, case["kind"])


def _check_field_vectors(eng):
    for case in VEC["field_vectors"]:
        fid = case["field"]
        p = FIELD_MODULUS[fid]
        g = lambda k: pack(p, [unhex(h) for h in case[k]])
        u, r = mont_bytes(p, unhex(case["u"])), mont_bytes(p, unhex(case["r"]))
        assert eng.cross_term(fid, g("az"), g("bz"), g("cz"), g("e"), u) == g("cross_term")
        assert eng.fold(fid, g("az"), g("bz"), r) == g("fold")
        assert eng.bind_top(fid, g("az")[:8 * 32], r) == g("bind_top_of_az8")


def _check_sumchecks(eng):
    for case in VEC["sumcheck"]:
        fid, l = case["field"], case["num_rounds"]
        p = FIELD_MODULUS[fid]
        rng = SplitMix64(case["seed"])
        n = 1 << l
        A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
        taus = [rng.field(p) for _ in range(l)]
        tr = Keccak256Transcript(p, case["transcript_label"].encode())
        if case["kind"] == "quad_prod":
            polys, rs, finals = eng.prove_quad_prod(fid, unhex(case["claim"]), l, A, B, tr)
        else:
            assert taus == [unhex(h) for h in case["taus"]]
            polys, rs, finals = eng.prove_cubic3(fid, unhex(case["claim"]), taus, A, B, C, tr)
        assert [[to_repr(x).hex() for x in q] for q in polys] == case["polys"], case["kind"]
        assert [to_repr(x).hex() for x in rs] == case["challenges"]
        assert [to_repr(x).hex() for x in finals] == case["finals"]
        assert to_repr(tr.squeeze(b"end")).hex() == case["squeeze_after"]


# ------------------------------------------------------------------ C restatement (CPU) --------
def test_c_oracle_reproduces_commit_vectors(oracle):
    _check_commits(OracleEngine(oracle))


def test_c_oracle_reproduces_field_and_sumcheck_vectors(oracle):
    _check_field_vectors(OracleEngine(oracle))
    _check_sumchecks(OracleEngine(oracle))


def test_vectors_are_reproducible(tmp_path):
    """make_vectors.py is deterministic: regenerating gives the committed file (spot-checked on the
    cheap sections so the CPU suite stays fast)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_vectors", os.path.join(HERE, "golden", "make_vectors.py"))
    mv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mv)
    assert mv.field_vector_cases() == VEC["field_vectors"]
    assert mv.sumcheck_cases() == VEC["sumcheck"]


# ------------------------------------------------------------------ CUDA path (GPU) ------------
@pytest.mark.gpu
def test_device_reproduces_commit_vectors(b200):
    _check_commits(DeviceEngine(b200))


@pytest.mark.gpu
def test_device_reproduces_field_and_sumcheck_vectors(b200):
    _check_field_vectors(DeviceEngine(b200))
    _check_sumchecks(DeviceEngine(b200))
