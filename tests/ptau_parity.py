"""Shared body of the PTAU tests (nova_b200/ptau.py): the reference's own four read_ptau tests
(src/provider/ptau.rs:486-548) restated, plus the header / truncation errors, the non-canonical coordinate,
save_setup -> load_setup round trip and a commitment through the loaded key.  Run on the emulated device by
tests/test_ptau_cpu.py and on a B200 by tests/test_zz_new_paths_gpu.py."""
import io
import struct

import pytest

from oracle.pyref import CURVES, SplitMix64, mont_bytes

Q = CURVES[0].p  # bn256 base field


# ---- test-side Fq2 helpers (independent of nova_b200.ptau's): sqrt for the non-subgroup point -------------
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_sqrt(a):
    """square root in Fq[u]/(u^2+1), q = 3 mod 4 (norm method); None if a is not a square"""
    if a == (0, 0):
        return (0, 0)
    n = (a[0] * a[0] + a[1] * a[1]) % Q
    s = pow(n, (Q + 1) // 4, Q)
    if s * s % Q != n:
        return None
    inv2 = pow(2, -1, Q)
    for sign in (1, -1):
        t = (a[0] + sign * s) * inv2 % Q
        x0 = pow(t, (Q + 1) // 4, Q)
        if x0 * x0 % Q != t or x0 == 0:
            continue
        x1 = a[1] * pow(2 * x0, -1, Q) % Q
        if f2_mul((x0, x1), (x0, x1)) == a:
            return (x0, x1)
    return None


def non_subgroup_g2(ptau):
    """ptau.rs:467-484: x = 1, 2, ... until x^3 + b is a square and the point is not torsion-free"""
    x = (1, 0)
    while True:
        rhs = ((f2_mul(f2_mul(x, x), x)[0] + ptau._B2[0]) % Q, (f2_mul(f2_mul(x, x), x)[1] + ptau._B2[1]) % Q)
        y = f2_sqrt(rhs)
        if y is not None:
            P = (x, y)
            assert ptau.g2_on_curve(P)
            if not ptau.g2_is_torsion_free(P):
                return P
        x = ((x[0] + 1) % Q, x[1])


def g1_raw(P):
    return CURVES[0].affine_bytes(P)


def build(ptau, g1_pts, g2_pts, power):
    buf = io.BytesIO()
    ptau.write_ptau(buf, b"".join(g1_raw(P) for P in g1_pts), b"".join(ptau.g2_to_raw(P) for P in g2_pts), power)
    return buf.getvalue()


def run_reference_cases(ptau):
    c = CURVES[0]
    G1, G2 = c.gen, ptau.G2_GENERATOR
    one_one_g2 = ((1, 0), (1, 0))
    # test_read_ptau_accepts_subgroup_g2 (ptau.rs:486-495)
    data = build(ptau, [G1, G1], [G2, G2], 1)
    r1, r2 = ptau.read_ptau(io.BytesIO(data), 2, 2)
    assert r1 == g1_raw(G1) * 2 and r2 == ptau.g2_to_raw(G2) * 2
    # test_read_ptau_rejects_non_subgroup_g2 (ptau.rs:497-508)
    data = build(ptau, [G1, G1], [G2, non_subgroup_g2(ptau)], 1)
    with pytest.raises(ptau.PointNotInSubgroup):
        ptau.read_ptau(io.BytesIO(data), 2, 2)
    # test_read_ptau_rejects_off_curve_g1 (ptau.rs:510-528): (1, 1)
    data = build(ptau, [(1, 1), G1], [G2, G2], 1)
    with pytest.raises(ptau.PointNotOnCurve):
        ptau.read_ptau(io.BytesIO(data), 2, 2)
    # test_read_ptau_rejects_off_curve_g2 (ptau.rs:530-548): (1, 1) in Fq2
    data = build(ptau, [G1, G1], [G2, one_one_g2], 1)
    with pytest.raises(ptau.PointNotOnCurve):
        ptau.read_ptau(io.BytesIO(data), 2, 2)
    # the same four through load_setup (register_checked: one upload, validated in place)
    ck = ptau.load_setup(io.BytesIO(build(ptau, [G1, G1], [G2, G2], 1)), None, 2)
    assert ck.n == 2 and ck.tau_H == ptau.g2_to_raw(G2)
    ck.release()
    for g1s, g2s, err in (([G1, G1], [G2, non_subgroup_g2(ptau)], ptau.PointNotInSubgroup),
                          ([(1, 1), G1], [G2, G2], ptau.PointNotOnCurve),
                          ([G1, G1], [G2, one_one_g2], ptau.PointNotOnCurve)):
        with pytest.raises(err):
            ptau.load_setup(io.BytesIO(build(ptau, g1s, g2s, 1)), None, 2)


def run_format_errors(ptau):
    c = CURVES[0]
    G1, G2 = c.gen, ptau.G2_GENERATOR
    good = build(ptau, [G1, G1], [G2, G2], 1)
    # byte layout of write_ptau (ptau.rs:170-269): magic, version, 11 sections; header section first
    assert good[:4] == b"ptau" and struct.unpack_from("<II", good, 4) == (1, 11)
    assert struct.unpack_from("<Iq", good, 12) == (1, 40) and struct.unpack_from("<I", good, 24) == (32,)
    assert int.from_bytes(good[28:60], "little") == Q and struct.unpack_from("<I", good, 60) == (1,)
    assert len(good) == 12 + 12 + 40 + 12 + 7 * 12 + 12 + 128 + 12 + 256
    with pytest.raises(ptau.InvalidHead):
        ptau.read_ptau(io.BytesIO(b"ptax" + good[4:]), 2, 2)
    with pytest.raises(ptau.UnsupportedVersion):
        ptau.read_ptau(io.BytesIO(good[:4] + struct.pack("<I", 2) + good[8:]), 2, 2)
    with pytest.raises(ptau.InvalidNumSections):
        ptau.read_ptau(io.BytesIO(good[:8] + struct.pack("<I", 5) + good[12:]), 2, 2)
    wrong_prime = bytearray(good)
    wrong_prime[28:60] = CURVES[0].q.to_bytes(32, "little")  # the scalar modulus instead of the base modulus
    with pytest.raises(ptau.InvalidPrime):
        ptau.read_ptau(io.BytesIO(bytes(wrong_prime)), 2, 2)
    # power 1: at most 2 G2 and 3 G1 points (ptau.rs:354-367)
    with pytest.raises(ptau.InsufficientPowerForG1) as e:
        ptau.read_ptau(io.BytesIO(good), 4, 2)
    assert (e.value.power, e.value.required) == (1, 3)
    with pytest.raises(ptau.InsufficientPowerForG2):
        ptau.read_ptau(io.BytesIO(good), 2, 3)
    with pytest.raises(ptau.IoError):  # truncated inside the G2 section
        ptau.read_ptau(io.BytesIO(good[:-10]), 2, 2)
    with pytest.raises(ptau.IoError):  # asks for more G1 points than the section holds: runs into the next section
        ptau.read_ptau(io.BytesIO(build(ptau, [G1, G1], [G2, G2], 4)[:12 + 12 + 40 + 12 + 84 + 12 + 128]), 3, 0)
    # a pruned file (3 sections: header, TauG1, TauG2; ptau.rs:285-292) is accepted
    pruned = (b"ptau" + struct.pack("<II", 1, 3) + struct.pack("<Iq", 1, 40) + struct.pack("<I", 32) +
              Q.to_bytes(32, "little") + struct.pack("<I", 1) + struct.pack("<Iq", 2, 128) + g1_raw(G1) * 2 +
              struct.pack("<Iq", 3, 256) + ptau.g2_to_raw(G2) * 2)
    assert ptau.read_ptau(io.BytesIO(pruned), 2, 2)[0] == g1_raw(G1) * 2
    # read_raw refuses a non-canonical coordinate (x + q is congruent to a valid x): io::Error, not PointNotOnCurve
    raw = g1_raw(G1)
    noncanon = (int.from_bytes(raw[:32], "little") + Q).to_bytes(32, "little") + raw[32:]
    buf = io.BytesIO()
    ptau.write_ptau(buf, raw + noncanon, ptau.g2_to_raw(G2) * 2, 1)
    with pytest.raises(ptau.IoError):
        ptau.read_ptau(io.BytesIO(buf.getvalue()), 2, 2)
    with pytest.raises(ptau.IoError):
        ptau.load_setup(io.BytesIO(buf.getvalue()), None, 2)
    g2raw = ptau.g2_to_raw(G2)
    g2bad = (int.from_bytes(g2raw[:32], "little") + Q).to_bytes(32, "little") + g2raw[32:]
    buf = io.BytesIO()
    ptau.write_ptau(buf, raw * 2, g2raw + g2bad, 1)
    with pytest.raises(ptau.IoError):
        ptau.read_ptau(io.BytesIO(buf.getvalue()), 2, 2)


def run_setup_round_trip(ptau, oracle, tmp_path):
    """a test SRS [tau^i]G1, tau_H = [tau]G2 (hyperkzg.rs:357-376) -> save_setup -> file -> load_setup (n not a power
    of two) -> the resident key commits like the in-memory key; a corrupted point deep in the file is named."""
    from nova_b200 import provider
    from oracle import hyperkzg_ref as hk
    cid, c = 0, CURVES[0]
    p = c.q
    rng = SplitMix64(99)
    tau = rng.field(p)
    n = 48  # load_setup reads n.next_power_of_two() = 64 points
    srs = hk.setup_srs(cid, 64, tau)
    h = oracle.gen_bases(cid, 1, 777)
    tau_H = ptau.g2_to_raw(ptau.g2_mul(ptau.G2_GENERATOR, tau))
    src = provider.CommitmentKey(provider.Curve.BN254_G1, srs, h)
    src.tau_H = tau_H
    path = tmp_path / "kzg_test.ptau"
    with open(path, "wb") as f:
        ptau.save_setup(src, f)
    ptau.check_sanity_of_ptau_file(path, 64, 2)
    with pytest.raises(ptau.InsufficientPowerForG1):  # power = log2(64) + 1 = 7 -> at most 255 G1 points
        ptau.check_sanity_of_ptau_file(path, 256, 2)
    with open(path, "rb") as f:
        ck = ptau.load_setup(f, h, n)
    assert (ck.n, ck.tau_H, ck.bases) == (64, tau_H, srs)
    v = [rng.field(p) for _ in range(n)]
    r = rng.field(p)
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    got = provider.CommitmentEngine(provider.Curve.BN254_G1).commit(ck, pack(v), mont_bytes(p, r))
    exp = c.add(c.msm_naive(v, [c.affine_from_bytes(srs[64 * i:64 * i + 64]) for i in range(n)]),
                c.mul(r, c.affine_from_bytes(h)))
    assert got == exp
    ck.release()
    src.release()
    # corrupt point 37 of the file's G1 section
    raw = bytearray(open(path, "rb").read())
    g1_off = raw.index(srs[:64])
    raw[g1_off + 64 * 37 + 40] ^= 0x10
    with pytest.raises(ptau.PtauFileError, match="37"):
        ptau.load_setup(io.BytesIO(bytes(raw)), h, n)
    # an invalid blinding generator is refused with the key
    bad_h = h[:32] + bytes(32)
    with pytest.raises(ptau.PointNotOnCurve, match="blinding"):
        ptau.load_setup(io.BytesIO(open(path, "rb").read()), bad_h, n)


def run_checked_registration(oracle, n, bad_lo, bad_hi):
    import ctypes
    from nova_b200 import native, provider
    cid = 0
    bases = oracle.gen_bases(cid, n + 1)
    L = native.lib()

    def reg(b, h):
        handle, bad = ctypes.c_uint64(0), ctypes.c_size_t(0)
        rc = L.b200_ck_register_checked(cid, provider._cbuf(b), n, provider._cbuf(h) if h else None, 0,
                                        ctypes.byref(handle), ctypes.byref(bad))
        return rc, handle.value, bad.value
    rc, handle, bad = reg(bases[:64 * n], bases[64 * n:])
    assert (rc, bad) == (0, ctypes.c_size_t(-1).value) and handle != 0
    sc = oracle.gen_scalars(CURVES[cid].scalar_field, 5, n)
    out = ctypes.create_string_buffer(96)
    native.check(L.b200_msm(handle, 0, provider._cbuf(sc), n, out))
    assert oracle.jacobian_to_affine(cid, out.raw) == oracle.msm(cid, sc, bases[:64 * n])
    native.check(L.b200_ck_release(handle))
    broken = bytearray(bases[:64 * n])
    broken[64 * bad_hi + 3] ^= 1
    broken[64 * bad_lo + 35] ^= 1
    rc, handle, bad = reg(bytes(broken), bases[64 * n:])
    assert (rc, handle, bad) == (7, 0, bad_lo)
    assert b"point" in L.b200_last_error()
    rc, handle, bad = reg(bases[:64 * n], bases[64 * n:64 * n + 32] + bytes(32))
    assert (rc, handle, bad) == (7, 0, n)


def run_pedersen_key_file(ptau, oracle, cid):
    """pedersen.rs:317-340, 383-393: save -> load (h first), commit through the loaded key, and the file errors"""
    from nova_b200 import provider
    c = CURVES[cid]
    p = c.q
    n = 24  # -> 32 bases + h
    pts = oracle.gen_bases(cid, 33, 4242)
    bases, h = pts[:64 * 32], pts[64 * 32:]
    src = provider.CommitmentKey(provider.Curve(cid), bases, h)
    buf = io.BytesIO()
    ptau.pedersen_save_setup(src, buf)
    src.release()
    data = buf.getvalue()
    assert data == b"PEDERSEN_KEY" + h + bases
    ck = ptau.pedersen_load_setup(io.BytesIO(data), n, provider.Curve(cid))
    assert (ck.n, ck.h, ck.bases) == (32, h, bases)
    rng = SplitMix64(31 + cid)
    v = [rng.field(p) for _ in range(n)]
    r = rng.field(p)
    got = provider.CommitmentEngine(provider.Curve(cid)).commit(ck, b"".join(mont_bytes(p, x) for x in v), mont_bytes(p, r))
    assert got == c.add(c.msm_naive(v, [c.affine_from_bytes(bases[64 * i:64 * i + 64]) for i in range(n)]),
                        c.mul(r, c.affine_from_bytes(h)))
    ck.release()
    with pytest.raises(ptau.InvalidHead):
        ptau.pedersen_load_setup(io.BytesIO(b"PEDERSEN_KEZ" + data[12:]), n, provider.Curve(cid))
    with pytest.raises(ptau.IoError):
        ptau.pedersen_load_setup(io.BytesIO(data[:-1]), n, provider.Curve(cid))
    with pytest.raises(ptau.IoError):  # a key file for fewer generators than asked for
        ptau.pedersen_load_setup(io.BytesIO(data), 33, provider.Curve(cid))
    bad = bytearray(data)
    bad[12 + 64 * 20 + 5] ^= 1  # point 20 of the file = base 19
    with pytest.raises(ptau.PointNotOnCurve, match="point 20 of the file"):
        ptau.pedersen_load_setup(io.BytesIO(bytes(bad)), n, provider.Curve(cid))
    bad = bytearray(data)
    bad[12 + 40] ^= 1  # h
    with pytest.raises(ptau.PointNotOnCurve, match="point 0"):
        ptau.pedersen_load_setup(io.BytesIO(bytes(bad)), n, provider.Curve(cid))
    # h non-canonical AND a base off the curve: the reference meets h first -> io error
    bad = bytearray(data)
    hx = int.from_bytes(h[:32], "little") + c.p
    bad[12:44] = hx.to_bytes(32, "little")
    bad[12 + 64 * 7 + 1] ^= 1
    with pytest.raises(ptau.IoError):
        ptau.pedersen_load_setup(io.BytesIO(bytes(bad)), n, provider.Curve(cid))
