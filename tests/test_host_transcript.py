"""nova_b200/csrc/transcript.cuh compiled for the HOST: the body of the device round kernel
(`k_sc_round`: round polynomial, Keccak transcript, challenge, claim / eq bound) drives complete
sum-check proofs and must reproduce the oracle's prover messages bit for bit.  Keccak-256 and
from_uniform are additionally pinned by the reference's literals (tests/golden/reference_kats.json)."""
import ctypes
import json
import os
import struct
import subprocess

import pytest

from oracle import pyref
from oracle.pyref import FIELD_MODULUS, Keccak256Transcript, SplitMix64, from_mont, keccak256, mont_bytes

HERE = os.path.dirname(os.path.abspath(__file__))
QUAD, CUBIC3_EQ, CUBIC3_EQ_M1 = 0, 1, 2


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    so = os.path.join(HERE, "hostcheck", "libhostcheck.so")
    csrc = os.path.join(HERE, "..", "nova_b200", "csrc")
    hdrs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def hc_simt():
    """The REAL kernel wrappers (k_sc_round, k_sc_round_batched) run as 32 host threads through the SIMT shim
    (tests/hostcheck/simt_host.h): lane roles, shared-memory hand-offs and barriers are exercised too."""
    src = os.path.join(HERE, "hostcheck", "simt_check.cpp")
    so = os.path.join(HERE, "hostcheck", "libhostcheck_simt.so")
    csrc = os.path.join(HERE, "..", "nova_b200", "csrc")
    deps = [src, os.path.join(HERE, "hostcheck", "simt_host.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(params=["sequential", "simt"])
def round_fn(request, hc, hc_simt):
    """The round step two ways: the kernel body called sequentially, and the real one-warp kernel on 32 threads."""
    return hc.hc_sc_round if request.param == "sequential" else hc_simt.hc_simt_sc_round


def _buf(b):
    return ctypes.create_string_buffer(b, len(b))


def test_keccak256_matches_reference_literal_and_oracle(hc):
    with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
        k = json.load(f)["keccak_example"]
    out = ctypes.create_string_buffer(32)
    data = bytes.fromhex(k["input_hex"])
    assert hc.hc_keccak256(_buf(data), ctypes.c_size_t(len(data)), out) == 0
    assert out.raw.hex() == k["digest_hex"]
    rng = SplitMix64(7)
    for n in (0, 1, 7, 8, 9, 135, 136, 137, 271, 272, 273, 500, 1087, 2000):
        data = rng.bytes(n)
        assert hc.hc_keccak256(_buf(data) if n else None, ctypes.c_size_t(n), out) == 0
        assert out.raw == keccak256(data), n


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_from_uniform(hc, fid):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(11 + fid)
    cases = [bytes(64), b"\xff" * 64, b"\xff" * 32 + bytes(32), bytes(32) + b"\xff" * 32,
             p.to_bytes(32, "little") + bytes(32), (p - 1).to_bytes(32, "little") + (p - 1).to_bytes(32, "little"),
             (((1 << 256) // p) * p).to_bytes(32, "little") + (((1 << 256) // p) * p + 1).to_bytes(32, "little")]
    cases += [rng.bytes(64) for _ in range(200)]
    out = ctypes.create_string_buffer(32)
    for c in cases:
        assert hc.hc_from_uniform(fid, _buf(c), out) == 0
        got = int.from_bytes(out.raw, "little")
        assert got < p and from_mont(p, got) == int.from_bytes(c, "little") % p


class HostRoundEngine:
    """State + one call per round, exactly what the device loop keeps in `b200_sc_state`."""

    def __init__(self, round_fn, fid, claim, transcript):
        self.round_fn, self.fid, self.p = round_fn, fid, FIELD_MODULUS[fid]
        self.state = _buf(mont_bytes(self.p, claim) + mont_bytes(self.p, 1) + struct.pack("<Q", transcript.round) +
                          transcript.state + struct.pack("<Q", 0))
        assert len(self.state.raw) == 144
        self.pending = transcript.buf
        self.tr = transcript

    def round(self, kind, res, tau=None):
        p = self.p
        resb = _buf(b"".join(mont_bytes(p, x) for x in res))
        taub = _buf(mont_bytes(p, tau)) if tau is not None else None
        tinv = _buf(mont_bytes(p, pow(tau, -1, p))) if tau else None
        poly, r = ctypes.create_string_buffer(96), ctypes.create_string_buffer(32)
        pend = _buf(self.pending) if self.pending else None
        assert self.round_fn(self.fid, kind, self.state, resb, taub, tinv, pend, len(self.pending), ord("p"),
                             ord("c"), poly, r) == 0
        self.pending = b""
        ncoef = 2 if kind == QUAD else 3
        coeffs = [int.from_bytes(poly.raw[32 * k:32 * k + 32], "little") for k in range(ncoef)]
        return coeffs, from_mont(p, int.from_bytes(r.raw, "little"))

    def finish(self):
        """Hands the transcript back to the host object (what the mirror does after the loop)."""
        raw = self.state.raw
        self.tr.round = struct.unpack("<Q", raw[64:72])[0]
        self.tr.state = raw[72:136]
        self.tr.buf = b""
        return from_mont(self.p, int.from_bytes(raw[:32], "little")), from_mont(self.p, int.from_bytes(raw[32:64], "little"))


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
@pytest.mark.parametrize("prefix_absorbs", [0, 3, 40])
def test_quad_prod_proof_through_device_round_code(round_fn, fid, prefix_absorbs):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(100 + fid + prefix_absorbs)
    l = 5
    A = [rng.field(p) for _ in range(1 << l)]
    B = [rng.field(p) for _ in range(1 << l)]
    claim = sum(a * b for a, b in zip(A, B)) % p
    t_ref, t_dev = Keccak256Transcript(p, b"hq"), Keccak256Transcript(p, b"hq")
    for k in range(prefix_absorbs):  # bytes absorbed before the sum-check starts (pending buffer)
        x = rng.field(p)
        t_ref.absorb_scalar(b"x", x)
        t_dev.absorb_scalar(b"x", x)
    exp_polys, exp_rs, exp_finals = pyref.prove_quad_prod(p, claim, l, A, B, t_ref)
    eng = HostRoundEngine(round_fn, fid, claim, t_dev)
    polys, rs = [], []
    for _ in range(l):
        h = len(A) // 2
        e0 = sum(A[i] * B[i] for i in range(h)) % p
        bc = sum((A[h + i] - A[i]) * (B[h + i] - B[i]) for i in range(h)) % p
        coeffs, r = eng.round(QUAD, [e0, bc])
        polys.append(coeffs)
        rs.append(r)
        A, B = pyref.bind_top(p, A, r), pyref.bind_top(p, B, r)
    assert polys == exp_polys and rs == exp_rs and [A[0], B[0]] == exp_finals
    final_claim, _ = eng.finish()
    assert final_claim == A[0] * B[0] % p
    assert t_dev.squeeze(b"next") == t_ref.squeeze(b"next")  # the transcript continues identically


@pytest.mark.parametrize("fid", [0, 3])
@pytest.mark.parametrize("zero_tau_at", [None, 0, 2, 5])
def test_cubic3_eq_proof_through_device_round_code(round_fn, fid, zero_tau_at):
    """Includes tau = 0 rounds, where the reference takes the third-sum fall-back (sumcheck.rs:696-698)."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(200 + fid)
    l = 6
    n = 1 << l
    A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
    taus = [rng.field(p) for _ in range(l)]
    if zero_tau_at is not None:
        taus[zero_tau_at] = 0
    eqt = pyref.eq_evals(p, taus)
    claim = sum(e * (a * b - c) for e, a, b, c in zip(eqt, A, B, C)) % p
    t_ref, t_dev = Keccak256Transcript(p, b"hc"), Keccak256Transcript(p, b"hc")
    t_ref.absorb_scalar(b"k", 9)
    t_dev.absorb_scalar(b"k", 9)
    exp_polys, exp_rs, exp_finals = pyref.prove_cubic_with_three_inputs(p, claim, taus, A, B, C, t_ref)
    eng = HostRoundEngine(round_fn, fid, claim, t_dev)
    eq = pyref.EqSumCheckInstance(p, taus)  # used for its table selection only
    polys, rs = [], []
    for j in range(l):
        h = len(A) // 2
        f = [eq.factor(i) for i in range(h)]
        t0 = sum((A[i] * B[i] - C[i]) * f[i] for i in range(h)) % p
        tinf = sum((A[h + i] - A[i]) * (B[h + i] - B[i]) * f[i] for i in range(h)) % p
        if taus[j] == 0:
            tm1 = sum(((2 * A[i] - A[h + i]) * (2 * B[i] - B[h + i]) - (2 * C[i] - C[h + i])) * f[i] for i in range(h)) % p
            coeffs, r = eng.round(CUBIC3_EQ_M1, [t0, tinf, tm1], taus[j])
        else:
            coeffs, r = eng.round(CUBIC3_EQ, [t0, tinf], taus[j])
        polys.append(coeffs)
        rs.append(r)
        A, B, C = (pyref.bind_top(p, Z, r) for Z in (A, B, C))
        eq.bound(r)
    assert polys == exp_polys and rs == exp_rs and [A[0], B[0], C[0]] == exp_finals
    _, q = eng.finish()
    assert q == eq.eval_eq_left
    assert t_dev.squeeze(b"next") == t_ref.squeeze(b"next")


def test_eq_tables_are_contiguous_tau_slices():
    """The per-round split-eq tables (sumcheck.rs:606-664) are eq tables of contiguous slices of tau:
    left[k] = eq(taus[fh-k : fh]), right[k] = eq(taus[l-k : l]) -- what the device loop builds with
    b200_eq_table_dev instead of uploading host-built tables."""
    p = FIELD_MODULUS[0]
    rng = SplitMix64(5)
    for l in (1, 2, 3, 6, 7):
        taus = [rng.field(p) for _ in range(l)]
        eq = pyref.EqSumCheckInstance(p, taus)
        fh = eq.first_half
        for k, tab in enumerate(eq.poly_eq_left):
            assert tab == pyref.eq_evals(p, taus[fh - k:fh])
        for k, tab in enumerate(eq.poly_eq_right):
            assert tab == pyref.eq_evals(p, taus[l - k:l])


def test_keccak_and_from_uniform_property_based(hc):
    """hypothesis: Keccak-256 of arbitrary messages up to the buffer limit and from_uniform of arbitrary 64 bytes."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=120, deadline=None)
    @given(st.binary(min_size=0, max_size=2175), st.integers(0, 3), st.binary(min_size=64, max_size=64))
    def prop(msg, fid, wide):
        out = ctypes.create_string_buffer(32)
        assert hc.hc_keccak256(_buf(msg) if msg else None, ctypes.c_size_t(len(msg)), out) == 0
        assert out.raw == keccak256(msg)
        p = FIELD_MODULUS[fid]
        o2 = ctypes.create_string_buffer(32)
        assert hc.hc_from_uniform(fid, _buf(wide), o2) == 0
        assert from_mont(p, int.from_bytes(o2.raw, "little")) == int.from_bytes(wide, "little") % p
    prop()
