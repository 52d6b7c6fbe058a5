"""Parity of the CUDA MSM path (through the C ABI) against the oracle.  Mirrors the reference's
own MSM tests: msm.rs:722-821, curve_property_tests.rs:172-218, blitzar.rs:48-214."""
import pytest

from oracle.pyref import CURVES, SplitMix64, mont_bytes

pytestmark = pytest.mark.gpu


def aff(c, b):
    return c.affine_from_bytes(b)


def make_key(b200, oracle, cid, n, h=False, window_bits=0):
    bases = oracle.gen_bases(cid, n + (1 if h else 0))
    ck = b200.CommitmentKey(b200.Curve(cid), bases[:64 * n], bases[64 * n:] if h else None, window_bits)
    return ck, bases


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [0, 1, 2, 8, 16, 17, 100, 1000, 8104, 8200])
def test_msm_random_scalars(b200, oracle, cid, n):
    """vartime_multiscalar_mul == oracle msm (== naive) on seeded uniform scalars; sizes from
    msm.rs:741 (8), curve_property_tests.rs:172-177 (16,100,8104,8200), blitzar.rs:48-120 (0,2,100)."""
    c = CURVES[cid]
    ck, bases = make_key(b200, oracle, cid, max(n, 1))
    sc = oracle.gen_scalars(c.scalar_field, 1000 * cid + n, n)
    got = b200.DlogGroup(cid).vartime_multiscalar_mul(sc, ck)
    assert got == aff(c, oracle.msm(cid, sc, bases[:64 * n]))


@pytest.mark.parametrize("cid", [0, 2])
@pytest.mark.parametrize("kind", ["equal", "alt", "zeros", "ones", "minus_ones", "small_mixed"])
def test_msm_structured_scalars(b200, oracle, cid, kind):
    """all-equal / alternating 0,(r-1) (curve_property_tests.rs:196-218) and the witness-like
    skews (0/1, -1, small signed) that drive the reference's partitioning (msm.rs:237-277)."""
    c = CURVES[cid]
    n = 8200
    ck, bases = make_key(b200, oracle, cid, n)
    rng = SplitMix64(17)
    q = c.q
    if kind == "equal":
        v = rng.field(q)
        vals = [v] * n
    elif kind == "alt":
        vals = [0 if i % 2 == 0 else q - 1 for i in range(n)]
    elif kind == "zeros":
        vals = [0] * n
    elif kind == "ones":
        vals = [rng.next() & 1 for _ in range(n)]
    elif kind == "minus_ones":
        vals = [q - 1] * n
    else:
        mags = [1, 3, 200, 60000, (1 << 31) + 5, (1 << 63) + 9, (1 << 100) + 3]
        vals = [(mags[i % 7] if (i // 7) % 2 == 0 else q - mags[i % 7]) for i in range(n)]
    sc = b"".join(mont_bytes(q, v) for v in vals)
    got = b200.DlogGroup(cid).vartime_multiscalar_mul(sc, ck)
    assert got == aff(c, oracle.msm(cid, sc, bases))


def test_msm_identity_bases_and_duplicates(b200, oracle):
    """identity bases with non-zero scalars are skipped (msm.rs:247, 788-811); duplicate bases
    force the P+P (doubling) and P+(-P) branches inside buckets (msm.rs:92-113,130-155)."""
    cid, c = 0, CURVES[0]
    n = 600
    bases = bytearray(oracle.gen_bases(cid, n))
    for i in (3, 77, 500):
        bases[64 * i:64 * i + 64] = bytes(64)
    # duplicates: same point many times with the same and with negated scalars
    P = bytes(bases[64 * 10:64 * 11])
    for i in range(100, 140):
        bases[64 * i:64 * i + 64] = P
    rng = SplitMix64(3)
    vals = [rng.field(c.q) for _ in range(n)]
    for i in range(100, 120):
        vals[i] = vals[100]
    for i in range(120, 140):
        vals[i] = c.q - vals[100]
    sc = b"".join(mont_bytes(c.q, v) for v in vals)
    ck = b200.CommitmentKey(b200.Curve(cid), bytes(bases))
    got = b200.DlogGroup(cid).vartime_multiscalar_mul(sc, ck)
    assert got == aff(c, oracle.msm(cid, sc, bytes(bases)))
    # all-identity key -> identity
    ck0 = b200.CommitmentKey(b200.Curve(cid), bytes(64 * 32))
    assert b200.DlogGroup(cid).vartime_multiscalar_mul(sc[:32 * 32], ck0) is None


@pytest.mark.parametrize("window_bits", [2, 4, 8, 11, 13, 16, 17, 20])
def test_msm_window_sizes(b200, oracle, window_bits):
    cid, c = 0, CURVES[0]
    n = 3000
    ck, bases = make_key(b200, oracle, cid, n, window_bits=window_bits)
    sc = oracle.gen_scalars(c.scalar_field, 5, n)
    assert b200.DlogGroup(cid).vartime_multiscalar_mul(sc, ck) == aff(c, oracle.msm(cid, sc, bases))


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_commit_with_blind(b200, oracle, cid):
    """CE::commit(ck, v, r) = MSM(v, ck[..len]) + h*r (pedersen.rs:263-270)."""
    c = CURVES[cid]
    n = 1024
    ck, bases = make_key(b200, oracle, cid, n, h=True)
    ce = b200.CommitmentEngine(cid)
    for m in (0, 1, 300, 1024):
        v = oracle.gen_scalars(c.scalar_field, 9 + m, m)
        r = oracle.gen_scalars(c.scalar_field, 99, 1)
        exp = aff(c, oracle.msm_naive(cid, v + r, bases[:64 * m] + bases[64 * n:]))
        assert ce.commit(ck, v, r) == exp
        assert ce.commit(ck, v, None) == aff(c, oracle.msm(cid, v, bases[:64 * m]))


def test_batch_ragged(b200, oracle):
    """batch_vartime_multiscalar_mul with varying lengths 0..100 (blitzar.rs:188-214)."""
    cid, c = 0, CURVES[0]
    ck, bases = make_key(b200, oracle, cid, 100)
    vecs = [oracle.gen_scalars(c.scalar_field, 7 * k + 1, k) for k in list(range(0, 100, 9)) + [100]]
    got = b200.DlogGroup(cid).batch_vartime_multiscalar_mul(vecs, ck)
    for v, g in zip(vecs, got):
        assert g == aff(c, oracle.msm(cid, v, bases[:2 * len(v)]))


@pytest.mark.parametrize("bits", [1, 4, 8, 10, 16, 20, 32, 40, 64])
def test_msm_small_widths(b200, oracle, bits):
    """msm_small == msm for the widths of msm.rs:751-774."""
    cid, c = 0, CURVES[0]
    n = 2000
    ck, bases = make_key(b200, oracle, cid, n)
    rng = SplitMix64(bits)
    vals = [rng.next() & ((1 << bits) - 1) for _ in range(n)]
    vals[0], vals[1] = 0, 1
    eb = 1 if bits <= 8 else 2 if bits <= 16 else 4 if bits <= 32 else 8
    got = b200.DlogGroup(cid).vartime_multiscalar_mul_small(vals, ck, elem_bytes=eb, max_num_bits=bits)
    assert got == aff(c, oracle.msm_small(cid, vals, bases, bits))


def test_batch_add_indices(b200, oracle):
    cid, c = 3, CURVES[3]
    n = 5000
    ck, bases = make_key(b200, oracle, cid, n)
    rng = SplitMix64(1)
    for m in (0, 1, 17, 4000):
        idx = [rng.next() % n for _ in range(m)]
        assert b200.DlogGroup(cid).batch_add(ck, idx) == aff(c, oracle.batch_add(cid, bases, idx))


def test_adhoc_msm(b200, oracle):
    """bases that are not a registered key (pedersen.rs:418-420)."""
    for cid in (0, 2):
        c = CURVES[cid]
        n = 700
        bases = oracle.gen_bases(cid, n, k0=12345)
        sc = oracle.gen_scalars(c.scalar_field, 4, n)
        assert b200.DlogGroup(cid).vartime_multiscalar_mul(sc, bases) == aff(c, oracle.msm(cid, sc, bases))


def test_full_size_closed_form(b200, oracle):
    """BASELINE config 2 size (2^20, BN254): with bases P_i = (k0+i)G the MSM must equal
    [sum_i s_i (k0+i)] G -- a size-independent check that needs no reference MSM."""
    cid, c = 0, CURVES[0]
    n = 1 << 20
    ck, bases = make_key(b200, oracle, cid, n)
    sc = oracle.gen_scalars(c.scalar_field, 2, n)
    got = b200.DlogGroup(cid).vartime_multiscalar_mul(sc, ck)
    k = oracle.dot_index(c.scalar_field, sc)
    assert got == aff(c, oracle.scalar_mul(cid, c.affine_bytes(c.gen), k))
    # linearity: MSM(s) + MSM(s') == MSM(s + s')
    sc2 = oracle.gen_scalars(c.scalar_field, 3, n)
    ssum = oracle.vec_add(c.scalar_field, sc, sc2)
    g2 = b200.DlogGroup(cid).vartime_multiscalar_mul(sc2, ck)
    g3 = b200.DlogGroup(cid).vartime_multiscalar_mul(ssum, ck)
    assert c.add(got, g2) == g3


def test_errors(b200, oracle):
    import nova_b200
    ck, _ = make_key(b200, oracle, 0, 16)
    with pytest.raises(AssertionError):  # msm.rs:226 assert_eq!(coeffs.len(), bases.len())
        b200.DlogGroup(0).vartime_multiscalar_mul(bytes(32 * 17), ck)
    from nova_b200.native import lib
    import ctypes
    out = ctypes.create_string_buffer(96)
    assert lib().b200_msm(ck.handle, 10, ctypes.create_string_buffer(32 * 16), 16, out) == 5  # B200_E_RANGE
    assert lib().b200_msm(123456, 0, None, 0, out) == 3  # B200_E_HANDLE
    ck.release()
    with pytest.raises(nova_b200.B200Error):
        b200.DlogGroup(0).vartime_multiscalar_mul(bytes(32), ck)


def test_config1_pedersen_commit_pallas_2p16(b200, oracle):
    """BASELINE.json configs[0]: Pedersen commit of 2^16 random Pallas scalars, bit-exact against the
    CPU restatement (commit = MSM(v, ck) + r*h, pedersen.rs:263-270; bench shape benches/commit.rs)."""
    cid, c = 2, CURVES[2]
    n = 1 << 16
    ck, bases = make_key(b200, oracle, cid, n, h=True)
    v = oracle.gen_scalars(c.scalar_field, 1, n)
    r = oracle.gen_scalars(c.scalar_field, 7, 1)
    ce = b200.CommitmentEngine(cid)
    assert ce.commit(ck, v, r) == aff(c, oracle.msm(cid, v + r, bases))
    assert ce.commit(ck, v, None) == aff(c, oracle.msm(cid, v, bases[:64 * n]))
    # commit is additively homomorphic: commit(v1) + commit(v2) == commit(v1 + v2)
    v2 = oracle.gen_scalars(c.scalar_field, 2, n)
    s = oracle.vec_add(c.scalar_field, v, v2)
    assert c.add(ce.commit(ck, v, None), ce.commit(ck, v2, None)) == ce.commit(ck, s, None)


def test_config4_size_2p22_closed_form(b200, oracle):
    """BASELINE.json configs[3] size (2^22 BN254 scalars): device-built key bases[i] = (k0+i)G, so the
    commitment must equal [sum_i s_i (k0+i)] G; plus prefix consistency commit(v[:n/2]) + commit of the
    zero-padded upper half == commit(v)."""
    cid, c = 0, CURVES[0]
    n = 1 << 22
    ck = b200.CommitmentKey.setup_synthetic(b200.Curve(cid), n, k0=oracle.K0_DEFAULT)
    sc = oracle.gen_scalars(c.scalar_field, 22, n)
    g = b200.DlogGroup(cid)
    got = g.vartime_multiscalar_mul(sc, ck)
    k = oracle.dot_index(c.scalar_field, sc)
    assert got == aff(c, oracle.scalar_mul(cid, c.affine_bytes(c.gen), k))
    half = n // 2
    lo = g.vartime_multiscalar_mul(sc[:32 * half], ck)
    hi = g.vartime_multiscalar_mul(bytes(32 * half) + sc[32 * half:], ck)
    assert c.add(lo, hi) == got


def test_short_vectors_on_a_wide_key(b200, oracle):
    """A key wide enough for 20-bit windows (n >= 2^22) routes MSMs that fit its first 2^21 bases
    to a second, 17-bit-window table set (capi.cu route()).  The result must not depend on which
    table set ran: compare with dedicated keys over the same bases, across the routing boundary,
    with and without the blinding term."""
    cid = 0
    c = CURVES[cid]
    big = b200.CommitmentKey.setup_synthetic(b200.Curve(cid), 1 << 22, with_h=True)
    ce = b200.CommitmentEngine(cid)
    sc = oracle.gen_scalars(c.scalar_field, 4242, (1 << 21) + 7 + 3)  # + 3: the vectors below start at element 3
    for m in (1, 1000, (1 << 16) + 3, 1 << 21, (1 << 21) + 7):
        own = b200.CommitmentKey.setup_synthetic(b200.Curve(cid), m)
        assert ce.commit(big, sc[:32 * m], None) == ce.commit(own, sc[:32 * m], None), m
        own.release()
    # oracle anchor on the narrow path
    bases = oracle.gen_bases(cid, 1000)
    assert ce.commit(big, sc[:32000], None) == aff(c, oracle.msm(cid, sc[:32000], bases))
    # r*h must come out the same through both table sets
    r = oracle.gen_scalars(c.scalar_field, 5, 1)
    zeros = bytes(32)
    assert ce.commit(big, zeros * 10, r) == ce.commit(big, zeros * ((1 << 21) + 5), r)
    # several vectors in one call, spread over the lanes of both table sets (b200_commit_many_dev)
    from nova_b200.spartan import DeviceVec, commit_many_dev
    lens = [(1 << 21) + 7, 1 << 21, 70000, 4096, 33, 2, 1, 1000, 5, (1 << 16) + 3]
    vecs = [DeviceVec.from_bytes(sc[32 * 3:32 * (3 + m)]) for m in lens]
    assert all(v.nbytes == 32 * m for v, m in zip(vecs, lens))  # (a short slice would make the device read past its end)
    many = commit_many_dev(cid, big, vecs, lens)
    for m, got in zip(lens, many):
        assert got == ce.commit(big, sc[32 * 3:32 * (3 + m)], None), m
    assert ce.batch_commit(big, [sc[:32 * m] for m in lens[2:]]) == [ce.commit(big, sc[:32 * m], None) for m in lens[2:]]
    # offset slices: inside the narrow range and straddling it
    g = b200.DlogGroup(cid)
    for off, m in ((12345, 5000), ((1 << 21) - 100, 300)):
        own = b200.CommitmentKey.setup_synthetic(b200.Curve(cid), m, k0=0x5EED + off)
        out = __import__("ctypes").create_string_buffer(96)
        from nova_b200.native import check, lib
        from nova_b200.provider import _cbuf, _jac_to_affine
        check(lib().b200_msm(big.handle, off, _cbuf(sc[:32 * m]), m, out))
        assert _jac_to_affine(b200.Curve(cid), out.raw) == ce.commit(own, sc[:32 * m], None)
        own.release()
    big.release()
