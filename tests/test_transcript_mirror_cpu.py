"""nova_b200.transcript (the product-side host mirror of src/provider/keccak.rs) against the reference's own literal
vectors (tests/golden/reference_kats.json: keccak.rs:222-258, 279-288) and against the oracle's independent
implementation on random inputs.  CPU only."""
import json
import os

from nova_b200.transcript import Keccak256Transcript, keccak256, keccak256_py
from oracle import pyref

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def test_keccak_example_digest():
    k = KATS["keccak_example"]
    assert keccak256(bytes.fromhex(k["input_hex"])).hex() == k["digest_hex"]
    assert keccak256_py(bytes.fromhex(k["input_hex"])).hex() == k["digest_hex"]


def test_transcript_challenges_from_the_reference():
    for case in KATS["keccak_transcript"]["cases"]:
        p = int(case["scalar_modulus_hex"], 16)
        t = Keccak256Transcript(p, b"test")
        t.absorb_scalar(b"s1", 2)
        t.absorb_scalar(b"s2", 5)
        assert t.squeeze(b"c1").to_bytes(32, "little").hex() == case["c1"], case["engine"]
        t.absorb_scalar(b"s3", 128)
        assert t.squeeze(b"c2").to_bytes(32, "little").hex() == case["c2"], case["engine"]


def test_equals_the_oracle_on_random_traffic():
    rng = pyref.SplitMix64(2024)
    for n in (0, 1, 7, 135, 136, 137, 271, 272, 273, 1000):
        d = rng.bytes(n)
        assert keccak256(d) == pyref.keccak256(d) == keccak256_py(d)
    p = pyref.FIELD_MODULUS[0]
    a, b = Keccak256Transcript(p, b"RelaxedR1CSSNARK"), pyref.Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    for k in range(6):
        blob = rng.bytes(50 * k)
        a.absorb_bytes(b"x", blob)
        b.absorb_bytes(b"x", blob)
        if k % 2:
            assert a.squeeze(b"c") == b.squeeze(b"c")
    assert (a.round, a.state, a.buf) == (b.round, b.state, b.buf)
