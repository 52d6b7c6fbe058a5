"""Host logic of nova_b200/ppsnark.py (the MicroSpartan prover core, ppsnark.rs:1056-1355) on the CPU through
tests/emulated_device.py: the body of the GPU parity test (tests/test_ppsnark_gpu.py) runs unchanged, so the
mirror's glue -- and the lifetime of every temporary device buffer, which the emulated device poisons on
free -- is checked without a GPU."""
import gc

import pytest

import emulated_device


@pytest.fixture()
def emulated():
    import nova_b200
    emulated_device.install()
    yield nova_b200
    gc.collect()
    emulated_device.uninstall()


@pytest.fixture()
def emulated_simt():
    """Same, with b200_sc_round_batched_dev answered by the REAL kernel k_sc_round_batched on 32 host threads."""
    import nova_b200
    emulated_device.install().use_simt = True
    yield nova_b200
    gc.collect()
    emulated_device.uninstall()


@pytest.mark.parametrize("cid,num_cons,num_vars", [(0, 8, 8), (1, 16, 8), (3, 4, 16)])
@pytest.mark.parametrize("device_transcript", [False, True])
def test_ppsnark_prove_core_host_logic(emulated, oracle, cid, num_cons, num_vars, device_transcript):
    """device_transcript=True: outer sum-check through the fused loop and the batched inner sum-check through
    b200_sc_round_batched_dev, both answered by the HOST BUILD of the device round kernels."""
    import test_ppsnark_gpu
    test_ppsnark_gpu.test_prove_core_matches_oracle(emulated, oracle, cid, num_cons, num_vars, device_transcript)


@pytest.mark.parametrize("zero_rho,zero_outer", [((), ()), ((0,), ()), ((2,), (1,)), ((0, 3), (0, 3))])
def test_batched_round_with_zero_taus(emulated, oracle, zero_rho, zero_outer, ell=4, helpers=None):
    """prove_helper with eq instances whose tau is 0 in some rounds (the third-sum fall-back,
    sumcheck.rs:1082-1213): host-transcript loop, device-transcript loop and the oracle's prove_helper agree.
    The engines are fed random polynomials directly (the prover does not check their consistency)."""
    from nova_b200 import ppsnark as dp
    from nova_b200 import spartan as sp
    from oracle import ppsnark_ref as pr
    from oracle.pyref import FIELD_MODULUS, Keccak256Transcript, SplitMix64, mont_bytes
    fid = 0
    p, N = FIELD_MODULUS[fid], 1 << ell
    rng = SplitMix64(5150 + len(zero_rho) + 7 * len(zero_outer))
    vec = lambda: [rng.field(p) for _ in range(N)]
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)
    oracles, aux, ts_row, ts_col = [vec() for _ in range(4)], [vec() for _ in range(4)], vec(), vec()
    L_row, L_col, val, E, W = vec(), vec(), vec(), vec(), vec()
    rhos = [0 if i in zero_rho else rng.field(p) for i in range(ell)]
    r_outer = [0 if i in zero_outer else rng.field(p) for i in range(ell)]
    claim, claim_E = rng.field(p), rng.field(p)
    num_vars = 4

    def ref_run():
        mem = pr.MemorySumcheckInstance(p, oracles, aux, rhos, ts_row, ts_col)
        inner = pr.InnerBatchedSumcheckInstance(p, claim, L_row, L_col, val, claim_E, r_outer, E)
        wit = pr.WitnessBoundSumcheck(p, r_outer, W, num_vars)
        tr = Keccak256Transcript(p, b"zt")
        tr.absorb_scalar(b"k", 3)
        return pr.prove_helper(p, mem, inner, wit, tr), tr.squeeze(b"after")

    def dev_run(helper):
        up = lambda v: sp.DeviceVec.from_bytes(pack(v))
        mem = dp.MemorySumcheckInstance(fid, N, [up(v) for v in oracles], [up(v) for v in aux], rhos, up(ts_row), up(ts_col))
        inner = dp.InnerBatchedSumcheckInstance(fid, N, claim, up(L_row), up(L_col), up(val), claim_E, r_outer, up(E))
        wit = dp.WitnessBoundSumcheck(fid, N, r_outer, up(W), num_vars)
        tr = Keccak256Transcript(p, b"zt")
        tr.absorb_scalar(b"k", 3)
        return helper(fid, mem, inner, wit, tr), tr.squeeze(b"after")

    exp, exp_after = ref_run()
    for helper in helpers or (dp.prove_helper, dp.prove_helper_device, dp.prove_helper_device_rounds):
        got, after = dev_run(helper)
        assert [list(q) for q in got[0]] == [list(q) for q in exp[0]], helper.__name__
        assert list(got[1]) == list(exp[1])
        assert got[2:] == exp[2:]
        assert after == exp_after


@pytest.mark.parametrize("zero_rho,zero_outer", [((), ()), ((2,), (1,)), ((0, 3), (0, 3))])
def test_batched_round_kernel_wrapper_on_32_threads(emulated_simt, oracle, zero_rho, zero_outer):
    """The one-warp kernel wrapper itself (lane i = claim i, lanes 0..2 combine, lanes 0/1 hash, shared-memory
    hand-offs between __syncwarp barriers) through the SIMT shim."""
    test_batched_round_with_zero_taus(emulated_simt, oracle, zero_rho, zero_outer)


def test_batched_sumcheck_unfused_round_on_host_threads(emulated_simt, oracle):
    """the four-launch form of a round (k_form_reduce_multi, k_form_final_multi, k_sc_round_batched; NOVA_B200_SC_UNFUSED=1)"""
    from nova_b200.native import lib
    lib().fused_round = False
    old = lib().b200_sumcheck_tail_bits(0)
    try:
        test_batched_round_with_zero_taus(emulated_simt, oracle, (0, 3), (1,))
    finally:
        lib().b200_sumcheck_tail_bits(old)


@pytest.mark.parametrize("tail_bits", [0, 4])
def test_batched_sumcheck_tail_kernel_on_host_threads(emulated_simt, oracle, tail_bits):
    """b200_sumcheck_batched with no tail (every round: k_form_reduce_multi, k_form_final_multi, k_sc_round_batched) and
    with ALL rounds inside k_scb_tail (512 host threads: block-wide sums, warp 0 runs the round incl. the pending
    transcript bytes of round 0, binds, barriers), against the oracle's prove_helper."""
    from nova_b200.native import lib
    assert lib().use_simt
    old = lib().b200_sumcheck_tail_bits(tail_bits)
    try:
        test_batched_round_with_zero_taus(emulated_simt, oracle, (0, 3), (1,))
    finally:
        lib().b200_sumcheck_tail_bits(old)


def test_ppsnark_prove_core_kernel_wrapper_on_32_threads(emulated_simt, oracle):
    import test_ppsnark_gpu
    test_ppsnark_gpu.test_prove_core_matches_oracle(emulated_simt, oracle, 0, 8, 8, True)


@pytest.mark.parametrize("num_cons,num_vars,device_transcript", [(8, 8, False), (8, 8, True)])
def test_whole_ppsnark_with_hyperkzg_host_logic(emulated, oracle, num_cons, num_vars, device_transcript):
    import ppsnark_full_parity
    ppsnark_full_parity.run(emulated, oracle, num_cons, num_vars, device_transcript)


@pytest.mark.parametrize("fid,l", [(0, 1), (0, 7), (3, 12)])
def test_eq_prefix_tables_kernel_on_host_threads(oracle, fid, l):
    """k_eq_prefix_tables (all nested eq tables of an EqSumCheckInstance in one launch) as 1024 host threads against the
    oracle's eq table of every slice taus[hi-k .. hi) (sumcheck.rs:606-664)."""
    import ctypes
    import emulated_device
    from oracle import coracle as co
    from oracle.pyref import FIELD_MODULUS, SplitMix64, mont_bytes
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(77 + l)
    taus = [rng.field(p) for _ in range(l)]
    if l > 2:
        taus[1] = 0
    raw = b"".join(mont_bytes(p, t) for t in taus)
    hc = emulated_device.EmulatedDevice()._hc_simt()
    for hi, K in ((l // 2, max(l // 2, 1) - 1), (l, l - l // 2)):
        out = ctypes.create_string_buffer(32 << (K + 1))
        assert hc.hc_simt_eq_prefix(fid, ctypes.create_string_buffer(raw, len(raw)), hi, K, out) == 0
        for k in range(K + 1):
            off = 32 * ((1 << k) - 1)
            assert out.raw[off:off + (32 << k)] == co.eq_table(fid, raw[32 * (hi - k):32 * hi]), (hi, k)


@pytest.mark.parametrize("fid", [0, 3])
def test_poly_eval_small_multi_kernel_on_host_threads(oracle, fid):
    """k_poly_eval_small_multi (the short tail of the HyperKZG fold chain at three points in one launch) as blocks of 256
    host threads against the oracle's Horner evaluation (hyperkzg.rs:1011-1019); lengths around the block size."""
    import ctypes
    import emulated_device
    from oracle import coracle as co
    from oracle.pyref import FIELD_MODULUS, SplitMix64, mont_bytes
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(4242 + fid)
    lens = [1, 2, 7, 255, 256, 257, 1000]
    polys = [b"".join(mont_bytes(p, rng.field(p)) for _ in range(n)) for n in lens]
    us = b"".join(mont_bytes(p, x) for x in (rng.field(p), 0, 1))
    bufs = [ctypes.create_string_buffer(b, len(b)) for b in polys]
    ptrs = (ctypes.c_void_p * len(lens))(*[ctypes.addressof(b) for b in bufs])
    out = ctypes.create_string_buffer(96 * len(lens))
    hc = emulated_device.EmulatedDevice()._hc_simt()
    assert hc.hc_simt_poly_eval_small_multi(fid, ptrs, (ctypes.c_size_t * len(lens))(*lens), len(lens),
                                            ctypes.create_string_buffer(us, 96), out) == 0
    for i, f in enumerate(polys):
        assert out.raw[96 * i:96 * i + 96] == co.poly_eval(fid, f, us), lens[i]
