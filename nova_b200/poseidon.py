"""Host mirror of the Poseidon random oracle (src/provider/poseidon.rs:18-127) with the squeeze on the device.

`PoseidonConstantsCircuit::default()` (poseidon.rs:27-36) = the neptune constants for the wide (arity 24) and narrow
(arity 5) sponges: `Sponge::api_constants(Strength::Standard)` -> `PoseidonConstants::new_with_strength_and_type(
Standard, HashType::Sponge)` (src/frontend/gadgets/poseidon/sponge/vanilla.rs:77-79, poseidon_inner.rs:147-216).
This module derives them the way the reference does -- round numbers from the security inequalities evaluated in f32
(round_numbers.rs:52-91), round constants from the Grain LFSR in self-shrinking mode (round_constants.rs:30-166), the
Cauchy MDS matrix 1 / (i + t + j) (mds.rs:104-134) -- registers them with the library once per (field, arity)
(b200_poseidon_register) and runs `squeeze` as one kernel (csrc/poseidon.cuh).

`PoseidonRO` has the reference's interface (absorb one base-field element at a time, squeeze(num_bits, start_with_one)
hashes everything absorbed so far and keeps only the hash as the new state).  `squeeze_dev` is the resident form: the
absorbed elements already sit in HBM (e.g. the coordinates of comm_T written by the commitment kernel) and the challenge
stays there, converted into the OTHER curve's field for the folds (base_as_scalar) -- no host round trip."""
from __future__ import annotations

import ctypes

import numpy as np

from . import fields
from .native import c_u64, check, lib

WIDE, NARROW = 24, 5
_f32 = np.float32


def _rounds_secure(t: int, rf: int, rp: int) -> bool:
    """round_numbers_are_secure (round_numbers.rs:76-91), every operation in f32 like the reference"""
    rpf, tf = _f32(rp), _f32(t)
    n, m = _f32(256.0), _f32(128.0)
    bounds = [_f32(6.0) if m <= (n - _f32(3.0)) * (tf + _f32(1.0)) else _f32(10.0),
              _f32(0.43) * m + _f32(np.log2(tf)) - rpf,
              _f32(0.21) * n - rpf,
              (_f32(0.14) * n - _f32(1.0) - rpf) / (tf - _f32(1.0))]
    return rf >= max(max(int(np.ceil(b)), 0) for b in bounds)  # `as usize` saturates at 0


def round_numbers(arity: int) -> tuple:
    """(R_F, R_P) of round_numbers_base: the cheapest secure pair with the security margin (+2 full rounds, +7.5 %
    partial rounds).  The reference mutates its outer loop variable inside the inner loop (round_numbers.rs:60-62):
    every further secure candidate of the same outer iteration carries two more full rounds -- reproduced here."""
    t = arity + 1
    best = None
    for rf0 in range(2, 1001, 2):
        rf = rf0
        for rp0 in range(4, 200):
            if not _rounds_secure(t, rf, rp0):
                continue
            rf += 2
            rp = int(np.ceil(_f32(1.075) * _f32(rp0)))
            cost = t * rf + rp
            if best is None or cost < best[0] or (cost == best[0] and rf < best[1]):
                best = (cost, rf, rp)
    return best[1], best[2]


def grain_constants(p: int, t: int, r_f: int, r_p: int) -> list:
    """generate_constants(field = 1, sbox = 1, n = NUM_BITS, t, R_F, R_P) (round_constants.rs:30-88): an 80-bit
    register (kept in one integer, bit 79 = the oldest bit), 160 warm-up steps, then bits in pairs -- emit the second of
    a pair whose first is 1 -- packed big-endian into NUM_BITS-bit integers, rejecting values >= p."""
    nbits = p.bit_length()
    reg = 0
    for width, val in ((2, 1), (4, 1), (12, nbits), (12, t), (10, r_f), (10, r_p), (30, (1 << 30) - 1)):
        reg = (reg << width) | val

    def step():
        nonlocal reg
        tap = lambda i: (reg >> (79 - i)) & 1  # b_i of the paper: index 0 = oldest
        new = tap(62) ^ tap(51) ^ tap(38) ^ tap(23) ^ tap(13) ^ tap(0)
        reg = ((reg << 1) & ((1 << 80) - 1)) | new
        return new
    for _ in range(160):
        step()

    def shrunk_bit():
        while True:
            first, second = step(), step()
            if first:
                return second
    out = []
    while len(out) < (r_f + r_p) * t:
        v = 0
        for _ in range(nbits):
            v = (v << 1) | shrunk_bit()
        if v < p:
            out.append(v)
    return out


def cauchy_mds(p: int, t: int) -> list:
    return [[pow(i + t + j, -1, p) for j in range(t)] for i in range(t)]


def io_pattern_tag(n_absorb: int, n_squeeze: int = 1, domain_separator: int = 0) -> int:
    """IOPattern([Absorb(n), Squeeze(m)]).value(ds) (sponge/api.rs:27-109), arithmetic mod 2^128; the library
    computes the same value inside b200_poseidon_ro* (this copy serves the tests)."""
    mask, base = (1 << 128) - 1, ((1 << 128) - 159)
    x_i = state = 0
    x_i = 1
    for a in ([n_absorb + (1 << 31)] if n_absorb else []) + ([n_squeeze] if n_squeeze else []) + [domain_separator]:
        x_i = (x_i * base) & mask
        state = (state + x_i * a) & mask
    return state


class PoseidonConstants:
    """One (field, arity) parameter set, resident on the device."""
    _cache: dict = {}

    def __init__(self, fid: int, arity: int):
        self.fid, self.arity, self.t = fid, arity, arity + 1
        self.p = fields.MODULUS[fid]
        self.r_f, self.r_p = round_numbers(arity)
        self.rc = grain_constants(self.p, self.t, self.r_f, self.r_p)
        self.mds = cauchy_mds(self.p, self.t)
        h = c_u64(0)
        rc = fields.pack(fid, self.rc)
        m = fields.pack(fid, [x for row in self.mds for x in row])
        check(lib().b200_poseidon_register(fid, arity, self.r_f, self.r_p, ctypes.create_string_buffer(rc, len(rc)),
                                           ctypes.create_string_buffer(m, len(m)), ctypes.byref(h)))
        self.handle = h.value

    @classmethod
    def get(cls, fid: int, arity: int = WIDE) -> "PoseidonConstants":
        key = (fid, arity)
        if key not in cls._cache:
            cls._cache[key] = cls(fid, arity)
        return cls._cache[key]


class PoseidonRO:
    """ROTrait for PoseidonRO (poseidon.rs:60-127) with the hash computed by the device kernel."""

    def __init__(self, fid: int, arity: int = WIDE):
        self.consts = PoseidonConstants.get(fid, arity)
        self.fid, self.p = fid, self.consts.p
        self.state: list = []

    def absorb(self, e: int):
        self.state.append(e % self.p)

    def squeeze(self, num_bits: int, start_with_one: bool = False) -> int:
        raw = fields.pack(self.fid, self.state)
        out = ctypes.create_string_buffer(96)
        check(lib().b200_poseidon_ro(self.consts.handle, ctypes.create_string_buffer(raw, max(len(raw), 1)),
                                     len(self.state), num_bits, int(start_with_one), out))
        h, c = fields.unpack(self.fid, out.raw[:64])
        assert int.from_bytes(out.raw[64:96], "little") == c
        self.state = [h]
        return c


def squeeze_dev(fid: int, d_elems, n: int, num_bits: int, out_field: int | None = None, arity: int = WIDE,
                start_with_one: bool = False):
    """Resident squeeze: `d_elems` = device vector of n Montgomery elements of field `fid`.  Returns a DeviceVec of 96
    bytes [hash | challenge | canonical challenge]; with `out_field` set, the second slot is rewritten as the challenge in
    THAT field's Montgomery form (what the folds of the other curve consume)."""
    from .spartan import DeviceVec
    consts = PoseidonConstants.get(fid, arity)
    out = DeviceVec(96)
    L = lib()
    check(L.b200_poseidon_ro_dev(consts.handle, d_elems.ptr, n, num_bits, int(start_with_one), out.ptr, None))
    if out_field is not None and out_field != fid:
        check(L.b200_to_mont_dev(out_field, ctypes.c_void_p(out.ptr.value + 64), 1, ctypes.c_void_p(out.ptr.value + 32), None))
    return out
