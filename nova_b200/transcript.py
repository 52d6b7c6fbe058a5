"""Host mirror of the reference's Fiat-Shamir transcript (src/provider/keccak.rs:31-160, the non-EVM variant):
Keccak-256 over `buffer || "NoDS" || round || state || label || {0,1}` with a 64-byte running state and challenges
taken by the wide reduction `from_uniform` (traits.rs TranscriptReprTrait / keccak.rs:139-160).

In the Rust integration the transcript stays in the host crate; this mirror exists so that the Python host layer can
drive the provers whose transcript steps sit between device calls (HyperKZG, ppsnark, NIFS) without any test-side
code.  The sum-check round loops carry their own copy of this logic on the device (csrc/transcript.cuh) and continue
the serialisable fields of this class (`round`, `state`, `buf`).  Pinned by the reference's golden vectors
(keccak.rs:241-258, 279-288) in tests/test_transcript_mirror_cpu.py."""
from __future__ import annotations

_MASK = (1 << 64) - 1
_ROUND_CONSTANTS = []
_ROTATION = [0] * 25
_PI_TARGET = [0] * 25  # lane i moves to _PI_TARGET[i] in the rho/pi step


def _init_tables():
    # iota constants from the degree-8 LFSR of the Keccak specification (x^8 + x^6 + x^5 + x^4 + 1)
    lfsr = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            if lfsr & 1:
                rc |= 1 << ((1 << j) - 1)
            lfsr = ((lfsr << 1) ^ (0x71 if lfsr & 0x80 else 0)) & 0xFF
        _ROUND_CONSTANTS.append(rc)
    # rho offsets / pi permutation: walk (x, y) -> (y, 2x + 3y) starting at (1, 0); offset t(t+1)/2
    x, y = 1, 0
    for t in range(24):
        _ROTATION[x + 5 * y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5
    for x in range(5):
        for y in range(5):
            _PI_TARGET[x + 5 * y] = y + 5 * ((2 * x + 3 * y) % 5)


_init_tables()


def _permute(s: list) -> list:
    """Keccak-f[1600] on 25 lanes, lane (x, y) at index x + 5 y."""
    for rc in _ROUND_CONSTANTS:
        col = [s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20] for x in range(5)]
        for x in range(5):
            c1 = col[(x + 1) % 5]
            d = col[(x + 4) % 5] ^ (((c1 << 1) | (c1 >> 63)) & _MASK)
            for y in range(0, 25, 5):
                s[x + y] ^= d
        b = [0] * 25
        for i in range(25):
            r = _ROTATION[i]
            v = s[i]
            b[_PI_TARGET[i]] = ((v << r) | (v >> (64 - r))) & _MASK if r else v
        for y in range(0, 25, 5):
            row = b[y:y + 5]
            for x in range(5):
                s[x + y] = row[x] ^ (~row[(x + 1) % 5] & _MASK & row[(x + 2) % 5])
        s[0] ^= rc
    return s


def keccak256(data: bytes) -> bytes:
    """Keccak-256 with the original 0x01 padding (sha3::Keccak256, keccak.rs:14): the library's host-side C
    implementation (b200_keccak256; no GPU involved).  `keccak256_py` below is the same function in the
    interpreter, kept as the independent cross-check of the C code (tests/test_transcript_mirror_cpu.py)."""
    import ctypes

    from .native import check, lib
    out = ctypes.create_string_buffer(32)
    check(lib().b200_keccak256(data, len(data), out))
    return out.raw


def keccak256_py(data: bytes) -> bytes:
    rate = 136
    padded = bytearray(data)
    padded.append(0x01)
    padded.extend(bytes(-len(padded) % rate))
    padded[-1] |= 0x80
    s = [0] * 25
    for off in range(0, len(padded), rate):
        for i in range(rate // 8):
            s[i] ^= int.from_bytes(padded[off + 8 * i:off + 8 * i + 8], "little")
        s = _permute(s)
    return b"".join(s[i].to_bytes(8, "little") for i in range(4))


class Keccak256Transcript:
    """`Keccak256Transcript::{new, absorb, squeeze}` (keccak.rs:98-160).  `p` is the modulus of the field the
    challenges live in (Engine::Scalar)."""

    PERSONA_TAG = b"NoTR"      # keccak.rs:22
    DOM_SEP_TAG = b"NoDS"      # keccak.rs:23

    def __init__(self, p: int, label: bytes):
        self.p = p
        self.round = 0
        self.buf = b""
        self.state = self._next_state(b"", self.PERSONA_TAG + label)

    @staticmethod
    def _next_state(buf: bytes, inp: bytes) -> bytes:
        # compute_updated_state (keccak.rs:66-95): the two hashes differ in one trailing counter byte
        pre = buf + inp
        return keccak256(pre + b"\x00") + keccak256(pre + b"\x01")

    def absorb_bytes(self, label: bytes, repr_: bytes):
        """absorb (keccak.rs:132-136): label, then the value's transcript representation"""
        self.buf += label + repr_

    def absorb_scalar(self, label: bytes, x: int):
        self.absorb_bytes(label, (x % self.p).to_bytes(32, "little"))

    def squeeze(self, label: bytes) -> int:
        """squeeze (keccak.rs:107-130): new state = H(buf || NoDS || round || state || label || b), challenge =
        from_uniform(new state) = the 512-bit little-endian integer reduced mod p"""
        out = self._next_state(self.buf, self.DOM_SEP_TAG + self.round.to_bytes(8, "little") + self.state + label)
        self.round += 1  # the reference uses a u16 with checked_add; a proof squeezes far fewer than 2^16 times
        self.state = out
        self.buf = b""
        return int.from_bytes(out, "little") % self.p
