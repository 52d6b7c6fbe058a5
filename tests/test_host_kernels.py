"""Whole CUDA kernels on the CPU: the two-stage reductions of every sum-check form (k_form_reduce over sc_form,
k_form_final; nova_b200/csrc/poly_kernels.cuh) run as <<<grid, 256>>> blocks of host threads through
tests/hostcheck/simt_host.h -- grid-stride loop, eq factor (split tables, shard mapping), warp shuffles, the
shared-memory stage and the final single-block pass execute as written -- and are compared with the C oracle.
(The arithmetic underneath is the host emulation of field.cuh, itself tested against Python integers and, chain
by chain, against the interpreted PTX.)"""
import ctypes
import os
import subprocess

import pytest

from oracle import coracle as co
from oracle.pyref import FIELD_MODULUS, SplitMix64, mont_bytes

HERE = os.path.dirname(os.path.abspath(__file__))
SC_NOUT = {0: 2, 1: 2, 2: 2, 3: 3, 4: 2, 5: 2, 6: 1, 7: 1, 8: 1, 9: 1, 10: 1, 11: 1}
EQ_FORMS = (4, 5, 6, 7, 8, 9, 10)


@pytest.fixture(scope="module")
def hc_simt():
    src = os.path.join(HERE, "hostcheck", "simt_check.cpp")
    so = os.path.join(HERE, "hostcheck", "libhostcheck_simt.so")
    csrc = os.path.join(HERE, "..", "nova_b200", "csrc")
    deps = [src, os.path.join(HERE, "hostcheck", "simt_host.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


def _buf(b):
    return (ctypes.c_char * max(len(b), 1)).from_buffer_copy(b) if b else None


def expected(fid, form, A, B, C, left, right, shift, id_mul=1, id_add=0):
    if form == 11:  # SC_DOT: plain inner product (provider/ipa_pc.rs:102-108); the C oracle has no form 11
        from oracle.pyref import from_mont_bytes
        p = FIELD_MODULUS[fid]
        n = len(A) // 32
        return mont_bytes(p, sum(from_mont_bytes(p, A[32 * i:32 * i + 32]) * from_mont_bytes(p, B[32 * i:32 * i + 32])
                                 for i in range(n)) % p)
    return co.sc_eval(fid, form, A, B, C, left, right, shift, id_mul, id_add)


def run_form(hc, fid, form, A, B, C, length, left, right, shift, id_mul, id_add, grid):
    out = ctypes.create_string_buffer(96)
    rc = hc.hc_simt_sc_eval(fid, form, _buf(A), _buf(B), _buf(C), ctypes.c_size_t(length), _buf(left), _buf(right),
                            shift, ctypes.c_size_t(id_mul), ctypes.c_size_t(id_add), grid, out)
    assert rc == 0
    return out.raw[:32 * SC_NOUT[form]]


@pytest.mark.parametrize("form", range(12))
@pytest.mark.parametrize("fid", [0, 3])
def test_sum_check_reduction_kernels_on_host_threads(hc_simt, fid, form):
    """length 2 * 600: 600 index pairs over 2 blocks of 256 threads, i.e. a grid-stride loop with a ragged last
    pass; eq factor split as left (8 entries) x right (128 entries, shift 7) and, for one case, unsplit."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(1000 * fid + form)
    length = 1200 if form not in (10, 11) else 777
    count = length if form in (10, 11) else length // 2
    vec = lambda n: b"".join(mont_bytes(p, rng.field(p)) for _ in range(n))
    A, B, C = vec(length), vec(length), vec(length)
    left, right, shift = vec(8), vec(128), 7
    assert count <= 8 * 128
    for (l, r, s) in ((left, right, shift), (None, vec(1024), 0)):
        if form not in EQ_FORMS and l is None:
            continue
        got = run_form(hc_simt, fid, form, A, B, C, length, l, r, s, 1, 0, 2)
        assert got == expected(fid, form, A, B, C, l, r, s)
    # one block, and more blocks than work (idle blocks contribute zero partials)
    assert run_form(hc_simt, fid, form, A, B, C, length, left, right, shift, 1, 0, 1) == \
        expected(fid, form, A, B, C, left, right, shift)
    if form in (4, 11):
        assert run_form(hc_simt, fid, form, A, B, C, length, left, right, shift, 1, 0, 5) == \
            expected(fid, form, A, B, C, left, right, shift)


@pytest.mark.parametrize("form", [4, 7, 10])
def test_sharded_index_mapping(hc_simt, form):
    """cyclic sharding: the local index j weighs with the eq factor of the global index j * id_mul + id_add"""
    fid, p = 0, FIELD_MODULUS[0]
    rng = SplitMix64(77 + form)
    vec = lambda n: b"".join(mont_bytes(p, rng.field(p)) for _ in range(n))
    length = 300 if form != 10 else 150
    A, B, C = vec(length), vec(length), vec(length)
    left, right, shift = vec(8), vec(64), 6  # 512 global indices >= 150 * 3
    for id_add in range(3):
        got = run_form(hc_simt, fid, form, A, B, C, length, left, right, shift, 3, id_add, 1)
        assert got == co.sc_eval(fid, form, A, B, C, left, right, shift, 3, id_add)


@pytest.mark.parametrize("fid,form", [(0, f) for f in EQ_FORMS] + [(3, 4), (3, 10)])
def test_segmented_eq_reduction_is_bit_identical(hc_simt, fid, form):
    """k_form_reduce_eqseg (opt-in NOVA_B200_SC_SEG=1): sum_hi left[hi] * sum_lo right[lo] X with pairs of indices
    sharing one reduction (fe_mul2_add) equals the flat sum -- with full segments, a ragged last segment, an odd
    number of indices per thread (the unpaired tail) and fewer blocks than segments."""
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(5000 + 10 * fid + form)
    vec = lambda n: b"".join(mont_bytes(p, rng.field(p)) for _ in range(n))
    shift = 10  # segments of 1024 indices = 4 per thread
    for count, grid in ((2 * 1024 + 700, 3), (2 * 1024 + 300, 2), (1024 + 1, 1)):
        length = count if form == 10 else 2 * count
        A, B, C = vec(length), vec(length), vec(length)
        left, right = vec(4), vec(1 << shift)
        out = ctypes.create_string_buffer(96)
        assert hc_simt.hc_simt_sc_eval_seg(fid, form, _buf(A), _buf(B), _buf(C), ctypes.c_size_t(length), _buf(left),
                                           _buf(right), shift, grid, out) == 0
        assert out.raw[:32 * SC_NOUT[form]] == co.sc_eval(fid, form, A, B, C, left, right, shift), (count, grid)
