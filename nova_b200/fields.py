"""Field / curve constants and the boundary encodings (host-side glue, big-int only for the
O(1) per-call conversions such as normalising one Jacobian result).

Moduli: src/provider/bn256_grumpkin.rs:39-40,84-85; src/provider/pasta.rs:37-38,45-46.
Encodings: 32-byte little-endian, Montgomery R = 2^256 in memory (halo2curves layout);
`to_repr` canonical bytes (src/provider/traits.rs:323-327).
"""
from __future__ import annotations

BN254_FR, BN254_FQ, PALLAS_FP, PALLAS_FQ = 0, 1, 2, 3
MODULUS = {
    BN254_FR: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    BN254_FQ: 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    PALLAS_FP: 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,
    PALLAS_FQ: 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001,
}
R = 1 << 256


def to_mont_bytes(fid: int, x: int) -> bytes:
    p = MODULUS[fid]
    return ((x % p) * R % p).to_bytes(32, "little")


def from_mont_bytes(fid: int, b: bytes) -> int:
    p = MODULUS[fid]
    return int.from_bytes(b, "little") * pow(R, -1, p) % p


def pack(fid: int, xs) -> bytes:
    return b"".join(to_mont_bytes(fid, x) for x in xs)


def unpack(fid: int, b: bytes):
    return [from_mont_bytes(fid, b[i:i + 32]) for i in range(0, len(b), 32)]
