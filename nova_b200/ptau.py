"""PTAU files <-> device-resident commitment keys ("PTAU loading onto device").

Host-side mirror of `src/provider/ptau.rs` (`read_ptau` :402-438, `write_ptau` :217-269, `read_meta_data`
:271-327, `read_header` :329-370, `check_sanity_of_ptau_file` :441-455) and of
`CommitmentEngine::{load_setup, save_setup}` for HyperKZG (hyperkzg.rs:657-689) and Pedersen
(pedersen.rs:317-340, 383-393), with the reference's error
type (`PtauFileError`, ptau.rs:100-151) as an exception hierarchy.

What moves to the device: the G1 section.  `write_raw` / `read_raw` (ptau.rs:197-208, 372-392) are the in-memory
Montgomery limbs of halo2curves, i.e. exactly the 64-byte base layout of include/nova_b200.h -- the section is
handed to `b200_ck_register_checked` unchanged, validated in HBM (canonical coordinates + curve equation, the two
checks of `read_points`) and expanded into the window tables only if every point passed.  The two G2 points of
the verifier key (`tau_H`) are O(1) control-plane data and are checked here on the host (curve equation over
Fq2 and [r]P = O, ptau.rs:424-435); `h = from_label(label, 1)` (hyperkzg.rs:672) stays with the caller.
"""
from __future__ import annotations

import ctypes
import io
import struct

from . import fields
from .native import c_size_t, c_u64, check, lib
from .provider import CommitmentKey, Curve, _cbuf

PTAU_VERSION = 1          # ptau.rs:161
NUM_SECTIONS_FULL = 11    # ptau.rs:163
NUM_SECTIONS_PRUNED = 3   # ptau.rs:165
MAX_PPOT_POWER = 28       # ptau.rs:168
B200_E_POINT = 7


class PtauFileError(Exception):
    """ptau.rs:100-151"""


class InvalidHead(PtauFileError):
    pass


class UnsupportedVersion(PtauFileError):
    def __init__(self, version):
        super().__init__(f"Unsupported version {version}")
        self.version = version


class InvalidNumSections(PtauFileError):
    def __init__(self, n):
        super().__init__(f"Invalid number of sections {n}")
        self.num_sections = n


class InvalidPrime(PtauFileError):
    def __init__(self, modulus):
        super().__init__(f"Invalid base prime {modulus:#x}")
        self.modulus = modulus


class InsufficientPowerForG1(PtauFileError):
    def __init__(self, power, required):
        super().__init__(f"Insufficient power for G1 (power {power}, required {required})")
        self.power, self.required = power, required


class InsufficientPowerForG2(PtauFileError):
    def __init__(self, power, required):
        super().__init__(f"Insufficient power for G2 (power {power}, required {required})")
        self.power, self.required = power, required


class PointNotOnCurve(PtauFileError):
    pass


class PointNotInSubgroup(PtauFileError):
    pass


class IoError(PtauFileError):
    """io::Error: short reads, and read_raw's refusal of a non-canonical coordinate"""


# ---------------------------------------------------------------------------------------------
# little helpers over a seekable binary reader (byteorder::ReadBytesExt, LittleEndian)
# ---------------------------------------------------------------------------------------------
def _read_exact(reader, n: int) -> bytes:
    b = reader.read(n)
    if b is None or len(b) != n:
        raise IoError("failed to fill whole buffer")
    return b


def _u32(reader) -> int:
    return struct.unpack("<I", _read_exact(reader, 4))[0]


def _i64(reader) -> int:
    return struct.unpack("<q", _read_exact(reader, 8))[0]


def read_meta_data(reader) -> dict:
    """ptau.rs:271-327 -> positions of sections 1 (header), 2 (TauG1), 3 (TauG2)"""
    try:
        if _read_exact(reader, 4).decode("utf-8") != "ptau":
            raise InvalidHead("Invalid magic string")
    except UnicodeDecodeError as e:
        raise PtauFileError(f"Utf8Error: {e}") from e
    version = _u32(reader)
    if version != PTAU_VERSION:
        raise UnsupportedVersion(version)
    num_sections = _u32(reader)
    if num_sections not in (NUM_SECTIONS_FULL, NUM_SECTIONS_PRUNED):
        raise InvalidNumSections(num_sections)
    pos = {1: 0, 2: 0, 3: 0}
    for _ in range(num_sections):
        sid = _u32(reader)
        size = _i64(reader)
        if sid in pos:
            pos[sid] = reader.tell()
        reader.seek(size, io.SEEK_CUR)
    if 0 in (pos[1], pos[2], pos[3]):  # the reference's assert_ne!s (ptau.rs:318-320): unconditional, also under -O
        raise InvalidNumSections(num_sections)
    return dict(pos_header=pos[1], pos_tau_g1=pos[2], pos_tau_g2=pos[3])


def read_header(reader, num_g1: int, num_g2: int, base_modulus: int) -> int:
    """ptau.rs:329-370; returns the power"""
    n8 = _u32(reader)
    modulus = int.from_bytes(_read_exact(reader, n8), "little")
    if modulus != base_modulus:
        raise InvalidPrime(modulus)
    power = _u32(reader)
    max_num_g2 = 1 << power
    max_num_g1 = max_num_g2 * 2 - 1
    if num_g1 > max_num_g1:
        raise InsufficientPowerForG1(power, max_num_g1)
    if num_g2 > max_num_g2:
        raise InsufficientPowerForG2(power, max_num_g2)
    return power


def check_sanity_of_ptau_file(path, num_g1: int, num_g2: int, curve: Curve = Curve.BN254_G1) -> None:
    """ptau.rs:441-455"""
    try:
        f = open(path, "rb")
    except OSError as e:
        raise IoError(str(e)) from e
    with f:
        meta = read_meta_data(f)
        f.seek(meta["pos_header"])
        read_header(f, num_g1, num_g2, fields.MODULUS[Curve(curve).base_field])


# ---------------------------------------------------------------------------------------------
# G2 of bn256 on the host: E'(Fq2): y^2 = x^3 + 3/(9+u), Fq2 = Fq[u]/(u^2+1).  O(1) points per key.
# ---------------------------------------------------------------------------------------------
_Q = fields.MODULUS[fields.BN254_FQ]
_R_ORDER = fields.MODULUS[fields.BN254_FR]
_INV82 = pow(82, -1, _Q)
_B2 = (27 * _INV82 % _Q, (-3 * _INV82) % _Q)  # 3 / (9 + u) = 3 (9 - u) / 82


def _f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % _Q, (a[0] * b[1] + a[1] * b[0]) % _Q)


def _f2_add(a, b):
    return ((a[0] + b[0]) % _Q, (a[1] + b[1]) % _Q)


def _f2_sub(a, b):
    return ((a[0] - b[0]) % _Q, (a[1] - b[1]) % _Q)


def _f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, _Q)
    return (a[0] * n % _Q, (-a[1]) * n % _Q)


def g2_on_curve(P) -> bool:
    """P = (x, y) with x, y in Fq2 as (c0, c1); None = identity (is_on_curve accepts it)"""
    if P is None:
        return True
    x, y = P
    return _f2_mul(y, y) == _f2_add(_f2_mul(_f2_mul(x, x), x), _B2)


def _g2_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    (x1, y1), (x2, y2) = P, Q
    if x1 == x2:
        if _f2_add(y1, y2) == (0, 0):
            return None
        xx = _f2_mul(x1, x1)
        lam = _f2_mul(_f2_add(_f2_add(xx, xx), xx), _f2_inv(_f2_add(y1, y1)))
    else:
        lam = _f2_mul(_f2_sub(y2, y1), _f2_inv(_f2_sub(x2, x1)))
    x3 = _f2_sub(_f2_sub(_f2_mul(lam, lam), x1), x2)
    return (x3, _f2_sub(_f2_mul(lam, _f2_sub(x1, x3)), y1))


def g2_mul(P, k: int):
    acc = None
    for bit in bin(k)[2:]:
        acc = _g2_add(acc, acc)
        if bit == "1":
            acc = _g2_add(acc, P)
    return acc


def g2_is_torsion_free(P) -> bool:
    """CofactorGroup::is_torsion_free (ptau.rs:431): [r]P = O"""
    return g2_mul(P, _R_ORDER) is None


def g2_from_raw(b: bytes):
    """128 bytes = x.c0, x.c1, y.c0, y.c1, each 32-byte little-endian Montgomery limbs (SerdeObject::read_raw);
    a coordinate >= q is read_raw's io::Error.  All-zero = identity."""
    cs = []
    for k in range(4):
        v = int.from_bytes(b[32 * k:32 * k + 32], "little")
        if v >= _Q:
            raise IoError("non-canonical G2 coordinate")
        cs.append(v * pow(fields.R, -1, _Q) % _Q)
    if cs == [0, 0, 0, 0]:
        return None
    return ((cs[0], cs[1]), (cs[2], cs[3]))


def g2_to_raw(P) -> bytes:
    if P is None:
        return bytes(128)
    (x, y) = P
    return b"".join(fields.to_mont_bytes(fields.BN254_FQ, c) for c in (x[0], x[1], y[0], y[1]))


# bn256 G2 generator (halo2curves bn256; the EIP-197 constants)
G2_GENERATOR = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def _check_g2(g2_raw: bytes, num_g2: int) -> None:
    """read_points::<G2> then the torsion loop, in the reference's order (ptau.rs:421-435): every point is
    parsed and checked on-curve first, subgroup membership afterwards."""
    pts = []
    for k in range(num_g2):
        P = g2_from_raw(g2_raw[128 * k:128 * k + 128])
        if not g2_on_curve(P):
            raise PointNotOnCurve("Point is not on the curve")
        pts.append(P)
    for P in pts:
        if not g2_is_torsion_free(P):
            raise PointNotInSubgroup("Point is not in the prime-order subgroup")


# ---------------------------------------------------------------------------------------------
# read / write
# ---------------------------------------------------------------------------------------------
def _read_sections(reader, num_g1: int, num_g2: int, curve: Curve):
    meta = read_meta_data(reader)
    reader.seek(meta["pos_header"])
    read_header(reader, num_g1, num_g2, fields.MODULUS[curve.base_field])
    reader.seek(meta["pos_tau_g1"])
    g1 = _read_exact(reader, 64 * num_g1)
    reader.seek(meta["pos_tau_g2"])
    g2 = _read_exact(reader, 128 * num_g2)
    return g1, g2


def _classify_bad_g1(curve: Curve, g1: bytes, idx: int) -> PtauFileError:
    """The device reports the first point that fails either check of read_points; which of the two it was
    (read_raw's canonicity -> io::Error, or is_on_curve -> PointNotOnCurve) is decided from its 64 bytes."""
    p = fields.MODULUS[curve.base_field]
    pt = g1[64 * idx:64 * idx + 64]
    if any(int.from_bytes(pt[k:k + 32], "little") >= p for k in (0, 32)):
        return IoError(f"non-canonical coordinate in G1 point {idx}")
    return PointNotOnCurve(f"Point is not on the curve (G1 point {idx})")


def read_ptau(reader, num_g1: int, num_g2: int, curve: Curve = Curve.BN254_G1):
    """ptau.rs:402-438 -> (g1 raw bytes, g2 raw bytes), every point validated: G1 on the device
    (`b200_ck_validate`), G2 on the host.  `load_setup` below is the form that leaves the key resident."""
    curve = Curve(curve)
    if num_g2 and curve != Curve.BN254_G1:
        raise ValueError("G2 sections are defined for bn256 only")
    g1, g2 = _read_sections(reader, num_g1, num_g2, curve)
    if num_g1:
        bad = CommitmentKey.validate(curve, g1)
        if bad is not None:
            raise _classify_bad_g1(curve, g1, bad)
    _check_g2(g2, num_g2)
    return g1, g2


def write_ptau(writer, g1_raw: bytes, g2_raw: bytes, power: int, curve: Curve = Curve.BN254_G1) -> None:
    """ptau.rs:217-269, byte for byte (the writer trusts its caller, ptau.rs:212-216)"""
    n8 = 32
    w = writer.write
    w(b"ptau")
    w(struct.pack("<II", PTAU_VERSION, NUM_SECTIONS_FULL))
    w(struct.pack("<Iq", 1, 4 + n8 + 4))
    w(struct.pack("<I", n8))
    w(fields.MODULUS[Curve(curve).base_field].to_bytes(n8, "little"))
    w(struct.pack("<I", power))
    w(struct.pack("<Iq", 0, 0))
    for sid in range(4, NUM_SECTIONS_FULL):
        w(struct.pack("<Iq", sid, 0))
    w(struct.pack("<Iq", 2, len(g1_raw)))
    w(g1_raw)
    w(struct.pack("<Iq", 3, len(g2_raw)))
    w(g2_raw)


def _next_power_of_two(n: int) -> int:
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def load_setup(reader, h: bytes | None, n: int, window_bits: int = 0) -> CommitmentKey:
    """`CommitmentEngine::load_setup` of HyperKZG (hyperkzg.rs:657-675): num = n.next_power_of_two() G1 points
    and 2 G2 points; ck = the G1 points, tau_H = the LAST G2 point, h = from_label(label, 1) -- derived by the
    caller and passed in (64 raw bytes, validated with the key; None for a key that never blinds).
    The G1 section goes to HBM once and is validated there; returns the resident key with `.tau_H` (128 raw
    bytes)."""
    curve = Curve.BN254_G1
    num = _next_power_of_two(n)
    g1, g2 = _read_sections(reader, num, 2, curve)
    handle, bad = c_u64(0), c_size_t(0)
    rc = lib().b200_ck_register_checked(int(curve), _cbuf(g1), num, _cbuf(h) if h else None, window_bits,
                                        ctypes.byref(handle), ctypes.byref(bad))
    if rc == B200_E_POINT:
        if h and bad.value == num:
            raise PointNotOnCurve("the blinding generator h is not a valid point")
        raise _classify_bad_g1(curve, g1, bad.value)
    check(rc)
    ck = CommitmentKey.from_handle(curve, handle.value, g1, h, num)
    try:
        _check_g2(g2, 2)
    except PtauFileError:
        ck.release()
        raise
    ck.tau_H = g2[128:256]
    return ck


def load_setup_sharded(reader, h: bytes | None, n: int, rank: int, world: int, group=None, window_bits: int = 0):
    """`load_setup` for one process per GPU (GPU g owns ck[g n/G .. (g+1) n/G) for the life of
    the key): every rank seeks to ITS index range of the TauG1 section and reads, uploads and validates only
    that slice -- file and PCIe traffic are divided by `world`, no rank ever holds the whole key.  The verdict
    is made global with one all-gather of (first bad global index, error class) so that every rank raises the
    same error the unsharded loader would (the smallest offending index decides, as in read_points' loop).
    Rank 0's slice carries the blinding generator h (the `r * h` term is added once).
    -> (resident key over the slice with `.tau_H`, lo, hi)"""
    from .sharding import all_gather_bytes, shard_range
    curve = Curve.BN254_G1
    num = _next_power_of_two(n)
    meta = read_meta_data(reader)
    reader.seek(meta["pos_header"])
    read_header(reader, num, 2, fields.MODULUS[curve.base_field])
    if world > num:  # decided identically on EVERY rank, before any rank-dependent branch: nobody is left in a collective
        raise ValueError(f"{world} ranks for a {num}-point key: some rank would own no point")
    lo, hi = shard_range(num, rank, world)
    reader.seek(meta["pos_tau_g1"] + 64 * lo)
    g1 = _read_exact(reader, 64 * (hi - lo))
    reader.seek(meta["pos_tau_g2"])
    g2 = _read_exact(reader, 256)
    my_h = h if rank == 0 else None
    handle, bad = c_u64(0), c_size_t(0)
    rc = lib().b200_ck_register_checked(int(curve), _cbuf(g1), hi - lo, _cbuf(my_h) if my_h else None, window_bits,
                                        ctypes.byref(handle), ctypes.byref(bad))
    NONE = (1 << 64) - 1
    verdict, kind = NONE, 0  # kind: 1 = PointNotOnCurve, 2 = non-canonical (io error), 3 = the blinding generator
    ck, lib_error = None, ""
    if rc == 0:
        ck = CommitmentKey.from_handle(curve, handle.value, g1, my_h, hi - lo)
    elif rc == B200_E_POINT:
        if my_h and bad.value == hi - lo:
            verdict, kind = num, 3  # after every G1 point, as in the unsharded order
        else:
            verdict = lo + bad.value
            kind = 2 if isinstance(_classify_bad_g1(curve, g1, bad.value), IoError) else 1
    else:  # a library / CUDA error on this rank: every rank must still reach the collective and stop
        lib_error = lib().b200_last_error().decode()
        verdict, kind = 0, 4
    try:
        votes = all_gather_bytes(verdict.to_bytes(8, "little") + bytes([kind]), group)
        if any(v[8] == 4 for v in votes):
            failed = [k for k, v in enumerate(votes) if v[8] == 4]
            raise IoError(f"key registration failed on rank(s) {failed}" + (f": {lib_error}" if lib_error else ""))
        first, first_kind = min((int.from_bytes(v[:8], "little"), v[8]) for v in votes)
        if first != NONE:
            if first_kind == 3:
                raise PointNotOnCurve("the blinding generator h is not a valid point")
            if first_kind == 2:
                raise IoError(f"non-canonical coordinate in G1 point {first}")
            raise PointNotOnCurve(f"Point is not on the curve (G1 point {first})")
        _check_g2(g2, 2)  # the same two points on every rank: the same verdict everywhere
    except Exception:
        if ck is not None:
            ck.release()
        raise
    ck.tau_H = g2[128:256]
    return ck, lo, hi


def sharded_commit(ck_slice: CommitmentKey, lo: int, hi: int, v_local: bytes, r: bytes | None, rank: int,
                   group=None):
    """`CE::commit(ck, v, r)` over a key loaded with `load_setup_sharded`: `v_local` = this rank's v[lo .. hi)
    clipped to len(v); every rank reduces its pairs to one point (rank 0 also adds r * h in the same pass), the
    96-byte partials are all-gathered and added.  -> affine (x, y) or None, on every rank."""
    import torch

    from .provider import _jac_to_affine
    from .sharding import all_gather_partials
    from .spartan import DeviceVec
    m = len(v_local) // 32
    assert m <= hi - lo
    out = ctypes.create_string_buffer(96)
    check(lib().b200_commit(ck_slice.handle, _cbuf(v_local), m, _cbuf(r) if (r and rank == 0) else None, out))
    parts = all_gather_partials(torch.frombuffer(bytearray(out.raw), dtype=torch.uint8), group)
    world = parts.numel() // 96
    d_parts, d_total = DeviceVec.from_bytes(parts.numpy().tobytes()), DeviceVec(96)
    check(lib().b200_jacobian_sum_dev(int(ck_slice.curve), d_parts.ptr, world, d_total.ptr, None))
    total = d_total.to_bytes()  # (the download synchronises; both buffers are still alive here)
    d_parts.free()
    d_total.free()
    return _jac_to_affine(ck_slice.curve, total)


# ---------------------------------------------------------------------------------------------
# Pedersen key files (provider/pedersen.rs:27-28, 317-340, 383-393): "PEDERSEN_KEY", then h, then the bases,
# all as raw points read with read_points
# ---------------------------------------------------------------------------------------------
KEY_FILE_HEAD = b"PEDERSEN_KEY"
_CURVE_B = {0: 3, 1: -17, 2: 5, 3: 5}  # bn256_grumpkin.rs:35-41,80-86; pasta.rs:33-47


def _host_point_error(curve: Curve, pt: bytes, what: str):
    """read_points' two checks for ONE point on the host (the blinding generator, which the file stores first)"""
    fid = curve.base_field
    p = fields.MODULUS[fid]
    if any(int.from_bytes(pt[k:k + 32], "little") >= p for k in (0, 32)):
        return IoError(f"non-canonical coordinate in {what}")
    x, y = fields.from_mont_bytes(fid, pt[:32]), fields.from_mont_bytes(fid, pt[32:])
    if (x, y) != (0, 0) and (y * y - x * x * x - _CURVE_B[int(curve)]) % p != 0:
        return PointNotOnCurve(f"Point is not on the curve ({what})")
    return None


def pedersen_load_setup(reader, n: int, curve: Curve, window_bits: int = 0) -> CommitmentKey:
    """`CommitmentEngine::load_setup` of Pedersen (pedersen.rs:317-340): head, then num + 1 points with
    num = n.next_power_of_two(); the FIRST point is h, the rest is ck.  The bases go to HBM once and are
    validated there."""
    curve = Curve(curve)
    num = _next_power_of_two(n)
    if _read_exact(reader, 12) != KEY_FILE_HEAD:
        raise InvalidHead("Invalid magic string")
    pts = _read_exact(reader, 64 * (num + 1))
    h, bases = pts[:64], pts[64:]
    err = _host_point_error(curve, h, "h, point 0 of the file")
    if err:
        raise err
    handle, bad = c_u64(0), c_size_t(0)
    rc = lib().b200_ck_register_checked(int(curve), _cbuf(bases), num, _cbuf(h), window_bits,
                                        ctypes.byref(handle), ctypes.byref(bad))
    if rc == B200_E_POINT:
        e = _classify_bad_g1(curve, bases, min(bad.value, num - 1))
        raise type(e)(f"{e} [point {bad.value + 1} of the file]")
    check(rc)
    return CommitmentKey.from_handle(curve, handle.value, bases, h, num)


def pedersen_save_setup(ck: CommitmentKey, writer) -> None:
    """pedersen.rs:383-393"""
    if ck.bases is None or ck.h is None:
        raise ValueError("this key has no host copy of its bases / blinding generator")
    writer.write(KEY_FILE_HEAD)
    writer.write(ck.h)
    writer.write(ck.bases)


def save_setup(ck: CommitmentKey, writer) -> None:
    """`save_setup` (hyperkzg.rs:677-689): g2 = [tau_H, tau_H], power = log2(next_power_of_two(len)) + 1"""
    if ck.bases is None:
        raise ValueError("this key was generated on the device and has no host copy of its bases")
    tau_h = getattr(ck, "tau_H", None)
    if tau_h is None:
        raise ValueError("key has no tau_H")
    power = (_next_power_of_two(ck.n).bit_length() - 1) + 1
    write_ptau(writer, ck.bases, tau_h + tau_h, power, ck.curve)

