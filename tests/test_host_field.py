"""The device field/curve templates (nova_b200/csrc/field.cuh, curve.cuh) compiled for the HOST,
where every PTX carry chain is replaced by its bit-exact emulation: validates the Montgomery
even/odd-accumulator algorithm and the XYZZ formulas without a GPU."""
import ctypes
import itertools
import os
import subprocess

import pytest

from oracle.pyref import CURVES, FIELD_MODULUS, SplitMix64, from_mont, mont_bytes, to_repr

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    so = os.path.join(HERE, "hostcheck", "libhostcheck.so")
    hdrs = [os.path.join(HERE, "..", "nova_b200", "csrc", f) for f in ("field.cuh", "curve.cuh", "field_constants.cuh", "field29.cuh", "curve29.cuh", "coop.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


def _buf(b):
    return ctypes.create_string_buffer(b, len(b))


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_field_ops(hc, fid):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(fid + 1)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (p - (1 << 256) % p) % p, (1 << 255) % p, (1 << 128) - 1, 1 << 64]
    pairs = list(itertools.product(edge, edge)) + [(rng.field(p), rng.field(p)) for _ in range(3000)]
    A = b"".join(mont_bytes(p, a) for a, _ in pairs)
    B = b"".join(mont_bytes(p, b) for _, b in pairs)
    n = len(pairs)
    for op, fn in [(0, lambda a, b: (a + b) % p), (1, lambda a, b: (a - b) % p), (2, lambda a, b: a * b % p), (7, lambda a, b: a * b % p),
                   (6, lambda a, b: (-a) % p)]:
        out = ctypes.create_string_buffer(32 * n)
        assert hc.hc_fe_op(fid, op, _buf(A), _buf(B), out, ctypes.c_size_t(n)) == 0
        for i, (a, b) in enumerate(pairs):
            got = int.from_bytes(out.raw[32 * i:32 * i + 32], "little")
            assert got < p
            assert from_mont(p, got) == fn(a, b), (op, a, b)
    m = 100  # inversion is slow on host; fewer cases
    out = ctypes.create_string_buffer(32 * m)
    hc.hc_fe_op(fid, 3, _buf(A[:32 * m]), _buf(A[:32 * m]), out, ctypes.c_size_t(m))
    for i, (a, _) in enumerate(pairs[:m]):
        assert from_mont(p, int.from_bytes(out.raw[32 * i:32 * i + 32], "little")) == (pow(a, -1, p) if a else 0)
    raw = b"".join(to_repr(a) for a, _ in pairs)
    out = ctypes.create_string_buffer(32 * n)
    hc.hc_fe_op(fid, 4, _buf(raw), _buf(raw), out, ctypes.c_size_t(n))
    assert out.raw == A
    out2 = ctypes.create_string_buffer(32 * n)
    hc.hc_fe_op(fid, 5, _buf(A), _buf(A), out2, ctypes.c_size_t(n))
    assert out2.raw == raw


@pytest.fixture(scope="module")
def hc_y3():
    """The same host build with the OTHER form of y3 in the mixed addition (two products, two reductions:
    -DNOVA_MADD_SPLIT_Y3); the default build computes y3 through fe_mul2_add (one reduction)."""
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    so = os.path.join(HERE, "hostcheck", "libhostcheck_splity3.so")
    csrc = os.path.join(HERE, "..", "nova_b200", "csrc")
    hdrs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-DNOVA_MADD_SPLIT_Y3", "-x", "c++",
                               src, "-o", so])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_xyzz_formulas_split_y3(hc_y3, cid):
    _xyzz_formulas(hc_y3, cid)
    # a long random walk of mixed additions: 3000 points, compared with the affine group law
    c = CURVES[cid]
    pts = c.bases_arith(3000, k0=0x1234 + cid)
    exp = None
    for P in pts:
        exp = c.add(exp, P)
    out = ctypes.create_string_buffer(96)
    hc_y3.hc_pt_sum(c.base_field, 0, _buf(b"".join(c.affine_bytes(P) for P in pts)), ctypes.c_size_t(len(pts)), out)
    assert c.jacobian_from_bytes(out.raw) == exp


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_xyzz_formulas(hc, cid):
    _xyzz_formulas(hc, cid)


def _xyzz_formulas(hc, cid):
    """madd / add / dbl incl. the exceptional cases of msm.rs:92-113,130-155."""
    c = CURVES[cid]
    pts = c.bases_arith(20)
    seq = [pts[3], pts[3], None, pts[5], c.neg(pts[5]), pts[7], pts[7], pts[7], pts[1], None] + pts
    exp = None
    for P in seq:
        exp = c.add(exp, P)
    data = b"".join(c.affine_bytes(P) for P in seq)
    for mode in (0, 1):
        out = ctypes.create_string_buffer(96)
        hc.hc_pt_sum(c.base_field, mode, _buf(data), ctypes.c_size_t(len(seq)), out)
        assert c.jacobian_from_bytes(out.raw) == exp
    d2 = c.affine_bytes(pts[2]) + c.affine_bytes(c.neg(pts[2]))
    out = ctypes.create_string_buffer(96)
    hc.hc_pt_sum(c.base_field, 0, _buf(d2), ctypes.c_size_t(2), out)
    assert c.jacobian_from_bytes(out.raw) is None
    for k in (1, 2, 3, 77, 65535, 1000003):
        out = ctypes.create_string_buffer(96)
        hc.hc_pt_sum(c.base_field, 2, _buf(c.affine_bytes(pts[4])), ctypes.c_size_t(k), out)
        assert c.jacobian_from_bytes(out.raw) == c.mul(k, pts[4])


@pytest.mark.parametrize("cid", [0, 2])
def test_quad_cooperative_ops(hc, cid):
    """coop.cuh: four lanes (here four std::threads with a barrier-based exchange) share one XYZZ
    addition / doubling and must all end with the same, correct point."""
    c = CURVES[cid]
    pts = c.bases_arith(12)
    seq = [pts[3], pts[3], None, pts[5], c.neg(pts[5]), pts[7], pts[7], pts[7], pts[1], None] + pts
    exp = None
    for P in seq:
        exp = c.add(exp, P)
    data = b"".join(c.affine_bytes(P) for P in seq)
    out = ctypes.create_string_buffer(96)
    hc.hc_coop(c.base_field, 0, _buf(data), ctypes.c_size_t(len(seq)), out)
    assert c.jacobian_from_bytes(out.raw) == exp
    for k in (1, 5, 17):
        out = ctypes.create_string_buffer(96)
        hc.hc_coop(c.base_field, 1, _buf(c.affine_bytes(pts[4])), ctypes.c_size_t(k), out)
        assert c.jacobian_from_bytes(out.raw) == c.mul(1 << k, pts[4])


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_on_curve_check(hc, cid):
    """affine_valid_raw (the device side of CommitmentKey::new's validation, hyperkzg.rs:113-119, and of
    read_points, ptau.rs:372-392): curve points and the identity encoding pass, any corrupted coordinate fails,
    and so does a coordinate that is congruent to a valid one but not canonical (x + p: read_raw refuses it)."""
    c = CURVES[cid]
    b_small = {0: 3, 1: -17, 2: 5, 3: 5}[cid]
    assert c.b % c.p == b_small % c.p
    pts = c.bases_arith(16)
    good = [c.affine_bytes(P) for P in pts] + [bytes(64)]
    bad = []
    for P in pts[:6]:
        bad.append(c.affine_bytes((P[0], (P[1] + 1) % c.p)))
        bad.append(c.affine_bytes(((P[0] + 1) % c.p, P[1])))
    bad.append(c.affine_bytes((0, 1)))
    n_plain_bad = len(bad)
    for k, P in enumerate(pts[6:10]):
        raw = c.affine_bytes(P)
        x, y = int.from_bytes(raw[:32], "little"), int.from_bytes(raw[32:], "little")
        if k & 1:
            bad.append((x + c.p).to_bytes(32, "little") + raw[32:])
        else:
            bad.append(raw[:32] + (y + c.p).to_bytes(32, "little"))
    data = b"".join(good + bad)
    ok = ctypes.create_string_buffer(len(good) + len(bad))
    assert hc.hc_on_curve(c.base_field, b_small, _buf(data), ctypes.c_size_t(len(good) + len(bad)), ok) == 0
    flags = list(ok.raw)
    assert flags[:len(good)] == [1] * len(good)
    exp_bad = [1 if c.on_curve(c.affine_from_bytes(x)) else 0 for x in bad[:n_plain_bad]]
    assert flags[len(good):len(good) + n_plain_bad] == exp_bad and sum(exp_bad) == 0
    assert flags[len(good) + n_plain_bad:] == [0] * 4  # non-canonical


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_sum_of_two_products_single_reduction(hc, fid):
    """fe_mul2_add(a, b, c, d) == a*b + c*d (one Montgomery reduction for both products), incl. the
    extreme operands p-1 that maximise every intermediate carry count."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(40 + fid)
    edge = [0, 1, p - 1, p - 2, (1 << 256) % p, (p - (1 << 256) % p) % p, (1 << 255) % p, (1 << 128) - 1]
    quads = list(itertools.product(edge, repeat=4)) + [tuple(rng.field(p) for _ in range(4)) for _ in range(4000)]
    # Montgomery-form limbs that are all-ones-ish: values whose REPRESENTATION is p-1 (largest limbs)
    big = from_mont(p, p - 1)
    quads += [(big, big, big, big), (big, big, 0, 0), (0, 0, big, big)]
    cols = [b"".join(mont_bytes(p, q[k]) for q in quads) for k in range(4)]
    n = len(quads)
    out = ctypes.create_string_buffer(32 * n)
    assert hc.hc_mul2_add(fid, _buf(cols[0]), _buf(cols[1]), _buf(cols[2]), _buf(cols[3]), out, ctypes.c_size_t(n)) == 0
    for i, (a, b, c, d) in enumerate(quads):
        got = int.from_bytes(out.raw[32 * i:32 * i + 32], "little")
        assert got < p and from_mont(p, got) == (a * b + c * d) % p, (a, b, c, d)


def test_field_identities_property_based(hc):
    """hypothesis: for random 256-bit inputs (reduced mod p) the host build of the device multiplier satisfies
    mul2_add(a, b, c, d) == mul(a, b) + mul(c, d) and distributes over addition, on all four fields."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(0, 3), st.lists(st.integers(0, (1 << 256) - 1), min_size=4, max_size=4))
    def prop(fid, xs):
        p = FIELD_MODULUS[fid]
        a, b, c, d = (x % p for x in xs)
        col = lambda v: _buf(mont_bytes(p, v))
        out = ctypes.create_string_buffer(32)
        assert hc.hc_mul2_add(fid, col(a), col(b), col(c), col(d), out, ctypes.c_size_t(1)) == 0
        assert from_mont(p, int.from_bytes(out.raw, "little")) == (a * b + c * d) % p
        m1, m2 = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        hc.hc_fe_op(fid, 2, col((a + c) % p), col(b), m1, ctypes.c_size_t(1))
        hc.hc_mul2_add(fid, col(a), col(b), col(c), col(b), m2, ctypes.c_size_t(1))
        assert m1.raw == m2.raw  # (a + c) b == a b + c b, bit for bit
    prop()


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_dedicated_squaring(hc, fid):
    """fe_sqr_dedicated (28 cross products + 8 squares + reduce-only rounds; A/B variant -DNOVA_SQR_DEDICATED)
    == a^2, incl. Montgomery REPRESENTATIONS made of all-ones limb patterns, which maximise every carry."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(60 + fid)
    reprs = [0, 1, p - 1, p - 2]
    for k in range(1, 254):
        reprs += [(1 << k) - 1, p - (1 << k), ((1 << k) - 1) << 1]
    reprs += [int("ffffffff" * w + "00000000" * (7 - w), 16) for w in range(8)]      # ones in the low limbs
    reprs += [int("7fffffff" + "ffffffff" * 6 + "fffffffe", 16) % p, int("55555555" * 8, 16) % p, int("aaaaaaaa" * 8, 16) % p]
    vals = [from_mont(p, r % p) for r in reprs] + [rng.field(p) for _ in range(3000)]
    A = b"".join(mont_bytes(p, a) for a in vals)
    out = ctypes.create_string_buffer(len(A))
    assert hc.hc_fe_op(fid, 8, _buf(A), _buf(A), out, ctypes.c_size_t(len(vals))) == 0
    for i, a in enumerate(vals):
        got = int.from_bytes(out.raw[32 * i:32 * i + 32], "little")
        assert got < p and from_mont(p, got) == a * a % p, hex(a)
