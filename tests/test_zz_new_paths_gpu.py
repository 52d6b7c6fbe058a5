"""GPU parity of the entry points and host flows added late in round 1.

This file sorts LAST on purpose: with `pytest -x` a failure here cannot hide the rest of the suite.

Ran on a B200 (60 passed in 10 s, profiles/r01s_new_paths_pytest.log) -- everything ABOVE the marker
"added after that run":
* sum-check round loops with the Keccak transcript on the device (b200_sumcheck_quad_prod, b200_sumcheck_cubic3)
* streamed witness hand-off (b200_witness_begin / append / finish)
* key validation (b200_ck_validate)

Written after the GPU budget ran out -- everything BELOW the marker.  Each of these has a CPU twin that runs the
same test body against the emulated device (tests/emulated_device.py, tests/cpp/emulated_b200.cpp), with the new
device code exercised through its host build (tests/hostcheck, incl. the real kernel wrappers on 32 threads):
* b200_witness_reset; the spartan::snark prover core and the whole SNARK with HyperKZG / IPA evaluation arguments;
  the device-resident folding step (commit_T, NIFS, folds, is_sat_relaxed, streamed witness, one half of
  CompressedSNARK::prove); the C++ mirror's resident layer; ppsnark with the batched device round
  (b200_sc_round_batched_dev); prove_batched_cubic; the remaining CommitmentEngine methods; the multi-GPU pieces
  beyond MSM and sum-check (two gloo ranks on one GPU);
* PTAU / Pedersen key files -> resident keys validated in HBM (b200_ck_register_checked, nova_b200/ptau.py), also
  one index range per rank; HyperKZG prove, the ppsnark batched sum-check and the folding step over two ranks;
  the segmented eq reductions (NOVA_B200_SC_SEG=1, in a subprocess).
"""
import os

import pytest

from oracle.pyref import (FIELD_MODULUS, Keccak256Transcript, SplitMix64, eq_evals, mont_bytes,
                          prove_cubic_with_three_inputs, prove_quad_prod)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def sp(b200):
    from nova_b200 import spartan
    return spartan


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


def _transcripts(p, label, absorbs, rng):
    a, b = Keccak256Transcript(p, label), Keccak256Transcript(p, label)
    for _ in range(absorbs):
        x = rng.field(p)
        a.absorb_scalar(b"x", x)
        b.absorb_scalar(b"x", x)
    return a, b


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
@pytest.mark.parametrize("l,absorbs", [(1, 0), (5, 2), (11, 0), (14, 40)])
def test_quad_prod_device_transcript(sp, fid, l, absorbs):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(500 + 7 * l + fid)
    A = [rng.field(p) for _ in range(1 << l)]
    B = [rng.field(p) for _ in range(1 << l)]
    claim = sum(a * b for a, b in zip(A, B)) % p
    t_ref, t_dev = _transcripts(p, b"dq", absorbs, rng)
    exp = prove_quad_prod(p, claim, l, A, B, t_ref)
    got = sp.SumcheckProof.prove_quad_prod_device(fid, claim, l, pack(p, A), pack(p, B), t_dev)
    assert got[1] == exp[1], "challenges"
    assert got[0] == exp[0], "compressed round polynomials"
    assert got[2] == exp[2], "final evaluations"
    assert t_dev.squeeze(b"n") == t_ref.squeeze(b"n"), "transcript state after the loop"


@pytest.mark.parametrize("fid", [0, 3])
@pytest.mark.parametrize("l,zero_tau_at", [(1, None), (2, None), (3, 1), (6, None), (6, 0), (7, 6), (12, None), (13, 4)])
def test_cubic3_device_transcript(sp, fid, l, zero_tau_at):
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(600 + 5 * l + fid)
    n = 1 << l
    A, B, C = ([rng.field(p) for _ in range(n)] for _ in range(3))
    taus = [rng.field(p) for _ in range(l)]
    if zero_tau_at is not None:
        taus[zero_tau_at] = 0
    eqt = eq_evals(p, taus)
    claim = sum(e * (a * b - c) for e, a, b, c in zip(eqt, A, B, C)) % p
    t_ref, t_dev = _transcripts(p, b"dc", 1, rng)
    exp = prove_cubic_with_three_inputs(p, claim, taus, A, B, C, t_ref)
    got = sp.SumcheckProof.prove_cubic_with_three_inputs_device(fid, claim, taus, pack(p, A), pack(p, B), pack(p, C),
                                                                t_dev)
    assert got[1] == exp[1], "challenges"
    assert got[0] == exp[0], "compressed round polynomials"
    assert got[2] == exp[2], "final evaluations"
    assert t_dev.squeeze(b"n") == t_ref.squeeze(b"n")


def test_device_loop_equals_host_loop_large(sp):
    """2^18 elements: the device-transcript loop against the existing host-transcript loop (itself
    parity-tested against the oracle at small sizes)."""
    fid, l = 0, 18
    p = FIELD_MODULUS[fid]
    from oracle import coracle as co
    A, B, C = (co.gen_scalars(fid, 70 + k, 1 << l) for k in range(3))
    rng = SplitMix64(77)
    taus = [rng.field(p) for _ in range(l)]
    claim = rng.field(p)  # the prover does not check the claim
    t1, t2 = Keccak256Transcript(p, b"big"), Keccak256Transcript(p, b"big")
    exp = sp.SumcheckProof.prove_cubic_with_three_inputs(fid, claim, taus, A, B, C, t1)
    got = sp.SumcheckProof.prove_cubic_with_three_inputs_device(fid, claim, taus, A, B, C, t2)
    assert got == exp
    assert t1.squeeze(b"n") == t2.squeeze(b"n")
    t1, t2 = Keccak256Transcript(p, b"big"), Keccak256Transcript(p, b"big")
    exp = sp.SumcheckProof.prove_quad_prod(fid, claim, l, A, B, t1)
    got = sp.SumcheckProof.prove_quad_prod_device(fid, claim, l, A, B, t2)
    assert got == exp


def test_pending_limit_is_an_error(sp):
    from nova_b200.native import B200Error
    fid = 0
    p = FIELD_MODULUS[fid]
    t = Keccak256Transcript(p, b"x")
    t.absorb_bytes(b"big", bytes(4000))
    with pytest.raises(B200Error):
        sp.SumcheckProof.prove_quad_prod_device(fid, 1, 1, pack(p, [1, 2]), pack(p, [3, 4]), t)


# ------------------------------------------------------------------ streamed witness hand-off ----
@pytest.mark.parametrize("cid", [0, 1, 2, 3])
@pytest.mark.parametrize("n,chunks", [(1, [1]), (37, [37]), (1000, [1, 2, 997]), (5000, [2500, 2500]),
                                       (70001, [4096] * 17 + [369])])
def test_witness_stream_commit_equals_whole_commit(b200, oracle, cid, n, chunks):
    from oracle.pyref import CURVES
    c = CURVES[cid]
    bases = oracle.gen_bases(cid, n + 1)
    ck = b200.CommitmentKey(b200.Curve(cid), bases[:64 * n], bases[64 * n:])
    v = oracle.gen_scalars(c.scalar_field, 900 + n, n)
    r = oracle.gen_scalars(c.scalar_field, 901, 1)
    exp = c.affine_from_bytes(oracle.msm(cid, v + r, bases))
    assert sum(chunks) == n
    ws = b200.WitnessStream(ck, n)
    off = 0
    for k in chunks:
        ws.append(v[32 * off:32 * (off + k)])
        off += k
    assert ws.finish(r) == exp
    assert ws.d_witness  # resident W is handed back
    from nova_b200.native import check, lib
    import ctypes
    back = ctypes.create_string_buffer(32 * n)
    check(lib().b200_memcpy_d2h(back, ws.d_witness, 32 * n))
    assert back.raw == v
    ws.release()
    assert b200.CommitmentEngine(cid).commit(ck, v, r) == exp  # the key's own workspace is untouched
    # no blind: r = 0
    ws = b200.WitnessStream(ck, n)
    ws.append(v)
    assert ws.finish(None) == c.affine_from_bytes(oracle.msm(cid, v, bases[:64 * n]))
    ws.release()
    ck.release()


def test_witness_stream_zero_extends_and_rejects_overflow(b200, oracle):
    """R1CSWitness::new_with_blind resizes the assignment to num_vars with zeros (r1cs/mod.rs:847-848)."""
    from oracle.pyref import CURVES
    from nova_b200.native import B200Error
    cid, n = 0, 3000
    c = CURVES[cid]
    bases = oracle.gen_bases(cid, n)
    ck = b200.CommitmentKey(b200.Curve(cid), bases)
    v = oracle.gen_scalars(c.scalar_field, 33, 2000)
    ws = b200.WitnessStream(ck, n)
    ws.append(v[:32 * 1500])
    ws.append(v[32 * 1500:])
    with pytest.raises(B200Error):
        ws.append(oracle.gen_scalars(c.scalar_field, 34, 1001))
    assert ws.finish(None) == c.affine_from_bytes(oracle.msm(cid, v, bases[:64 * 2000]))
    with pytest.raises(B200Error):
        ws.finish(None)
    ws.release()
    with pytest.raises(B200Error):
        b200.WitnessStream(ck, n + 1)
    ck.release()


def test_witness_stream_sparse_bits_witness(b200, oracle):
    """A sha256-like witness: mostly 0/1 with a few full-width values, pushed in 8 chunks, on a key
    large enough for the wide-window tables (2^19 -> 17-bit windows)."""
    from oracle.pyref import CURVES, mont_bytes as mb
    cid, n = 0, 1 << 19
    c = CURVES[cid]
    ck = b200.CommitmentKey.setup_synthetic(b200.Curve(cid), n)
    rng = SplitMix64(4242)
    one, zero = mb(c.q, 1), bytes(32)
    parts = []
    for i in range(n):
        x = rng.next()
        parts.append(mb(c.q, rng.field(c.q)) if x % 64 == 0 else (one if x & 1 else zero))
    v = b"".join(parts)
    ws = b200.WitnessStream(ck, n)
    step = n // 8
    for k in range(8):
        ws.append(v[32 * k * step:32 * (k + 1) * step])
    got = ws.finish(None)
    ws.release()
    assert got == b200.CommitmentEngine(cid).commit(ck, v, None)
    # closed form against the synthetic key P_i = (k0 + i) G:  sum s_i (k0 + i) G
    assert got == c.mul(oracle.dot_index(c.scalar_field, v), c.gen)
    ck.release()


# ------------------------------------------------------------------ key validation ---------------
@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_ck_validate(b200, oracle, cid):
    from oracle.pyref import CURVES
    c = CURVES[cid]
    n = 5000
    bases = bytearray(oracle.gen_bases(cid, n))
    assert b200.CommitmentKey.validate(b200.Curve(cid), bytes(bases)) is None
    bases[64 * 77:64 * 78] = bytes(64)  # the identity encoding is accepted (halo2curves is_on_curve)
    assert b200.CommitmentKey.validate(b200.Curve(cid), bytes(bases)) is None
    x, y = c.affine_from_bytes(bytes(bases[64 * 4321:64 * 4322]))
    bases[64 * 4321:64 * 4322] = c.affine_bytes((x, (y + 1) % c.p))
    assert b200.CommitmentKey.validate(b200.Curve(cid), bytes(bases)) == 4321
    bases[64 * 123:64 * 124] = c.affine_bytes(((x + 1) % c.p, y))
    assert b200.CommitmentKey.validate(b200.Curve(cid), bytes(bases)) == 123


# ================================================================== added after that run =========
def test_witness_stream_reset_reuse(b200, oracle):
    """One stream object serves successive prove_steps (b200_witness_reset keeps the workspace)."""
    from oracle.pyref import CURVES
    cid, n = 0, 9000
    c = CURVES[cid]
    bases = oracle.gen_bases(cid, n + 1)
    ck = b200.CommitmentKey(b200.Curve(cid), bases[:64 * n], bases[64 * n:])
    ws = b200.WitnessStream(ck, n)
    for step in range(3):
        v = oracle.gen_scalars(c.scalar_field, 40 + step, n)
        r = oracle.gen_scalars(c.scalar_field, 50 + step, 1) if step != 1 else None
        for off in range(0, n, 2048):
            ws.append(v[32 * off:32 * min(off + 2048, n)])
        exp = c.affine_from_bytes(oracle.msm(cid, v + r, bases) if r else oracle.msm(cid, v, bases[:64 * n]))
        assert ws.finish(r) == exp, step
        ws.reset()
    ws.release()
    ck.release()


@pytest.mark.parametrize("device_transcript", [True, False])
@pytest.mark.parametrize("cid,num_cons,num_vars,num_io", [(0, 4, 4, 1), (0, 16, 8, 2), (1, 8, 16, 2), (3, 32, 32, 3),
                                                          (0, 256, 128, 2)])
def test_snark_prove_core_matches_oracle(b200, oracle, cid, num_cons, num_vars, num_io, device_transcript):
    """spartan::snark::RelaxedR1CSSNARK::prove up to EE::prove (snark.rs:113-256) on the
    device; the same check runs on the CPU against an emulated device (tests/test_snark_mirror_cpu.py)."""
    from snark_parity import run_case
    run_case(b200, oracle, cid, num_cons, num_vars, num_io, device_transcript)


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_r1cs_fold_fixture_on_device(b200, oracle, cid):
    """The reference's folding fixture (nifs.rs:299-351: tiny cubic R1CS folded twice, is_sat_relaxed, then a
    relaxed + relaxed fold) through nova_b200.r1cs with everything resident; the same body runs on the CPU
    against the emulated device (tests/test_r1cs_mirror_cpu.py)."""
    from r1cs_parity import run_tiny_fixture
    run_tiny_fixture(b200, oracle, cid)


def test_cpp_mirror_resident_folding_step(oracle, tmp_path):
    """include/nova_b200.hpp's device-resident layer (DeviceVec, WitnessStream, validate_key,
    R1CSShapeDev::commit_T, fold_witness_resident) through tests/cpp/host_mirror_test --fold on the GPU; the
    same check runs on the CPU against the emulated library (tests/test_cpp_mirror.py)."""
    import test_cpp_mirror as tcm
    tcm.build()
    tcm.check_fold(tcm.EXE, oracle, tmp_path)


@pytest.mark.parametrize("cid,num_cons,num_vars", [(0, 8, 8), (1, 16, 8), (3, 4, 16), (0, 64, 64)])
def test_ppsnark_prove_core_device_transcript(b200, oracle, cid, num_cons, num_vars):
    """The MicroSpartan prover core with the outer sum-check as one fused call and the batched inner sum-check
    (prove_helper, ppsnark.rs:886-983) through b200_sc_round_batched_dev: every prover message equals the
    oracle's.  CPU twin (host build of the kernels + emulated device): tests/test_ppsnark_mirror_cpu.py."""
    import test_ppsnark_gpu
    test_ppsnark_gpu.test_prove_core_matches_oracle(b200, oracle, cid, num_cons, num_vars, True)


@pytest.mark.parametrize("zero_rho,zero_outer", [((), ()), ((0,), ()), ((2,), (1,)), ((0, 3), (0, 3))])
def test_batched_round_with_zero_taus_on_device(b200, oracle, zero_rho, zero_outer):
    import test_ppsnark_mirror_cpu as t
    t.test_batched_round_with_zero_taus(b200, oracle, zero_rho, zero_outer)


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_commit_variants_on_device(b200, oracle, cid):
    """commit_small_range, commit_sparse_binary, commit_sparse with and without the blind (pedersen.rs:285-305,
    396-427) through the provider mirror; CPU twin: tests/test_provider_mirror_cpu.py."""
    import commit_variants_parity
    commit_variants_parity.run(b200, oracle, cid)


@pytest.mark.parametrize("fid", [0, 3])
@pytest.mark.parametrize("k,l,zero", [(1, 3, ()), (3, 9, ()), (4, 6, (0, 3)), (16, 2, (1,)), (2, 13, ())])
def test_prove_batched_cubic_on_device(sp, fid, k, l, zero):
    """SumcheckProof::prove_batched_cubic (sumcheck.rs:513-577) through the mirror; CPU twin in
    tests/test_spartan_mirror_cpu.py."""
    import batched_cubic_parity
    batched_cubic_parity.run(sp, fid, k, l, zero)


def test_sharded_pieces_two_ranks_on_device(tmp_path):
    """SURVEY §8e pieces beyond the MSM and the sum-check -- SpMV / cross term on row slices, Horner evaluation
    and division by (X - u) on index-range slices -- two gloo ranks, both driving cuda:0 through
    nova_b200.sharding.DeviceEngine.  CPU twins: tests/test_sharding_pieces.py."""
    import test_sharding_pieces as t
    t.run_world(2, "gpu", tmp_path)


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_streamed_witness_folding_steps_on_device(b200, oracle, cid):
    """prove_step-shaped folds: WitnessStream (re-armed every step) -> resident W2 -> NIFS::prove, everything
    resident; CPU twin in tests/test_r1cs_mirror_cpu.py."""
    from r1cs_parity import run_streamed_steps
    run_streamed_steps(b200, oracle, cid)


@pytest.mark.parametrize("ell,tail_bits", [(6, 0), (6, 3), (9, 0), (9, 5), (9, 9), (11, 8)])
def test_batched_sumcheck_one_call_multi_and_tail(b200, oracle, ell, tail_bits):
    """b200_sumcheck_batched at sizes where the reductions span several blocks, with the hand-over to k_scb_tail at
    different rounds (0 = no tail: k_form_reduce_multi every round; = ell: every round inside the tail kernel), zero
    taus on both sides of the hand-over; against the oracle's prove_helper (proof, challenges, final claims, transcript)."""
    import test_ppsnark_mirror_cpu as t
    from nova_b200 import ppsnark as dp
    from nova_b200.native import lib
    old = lib().b200_sumcheck_tail_bits(tail_bits)
    try:
        t.test_batched_round_with_zero_taus(b200, oracle, (1, ell - 2), (0, ell - 1), ell=ell, helpers=(dp.prove_helper_device,))
    finally:
        lib().b200_sumcheck_tail_bits(old)


@pytest.mark.parametrize("cid,num_cons,num_vars,device_transcript", [(0, 16, 8, False), (0, 64, 64, True), (1, 32, 16, True)])
def test_compressed_snark_half_on_device(b200, oracle, cid, num_cons, num_vars, device_transcript):
    """One curve's half of CompressedSNARK::prove (nova/mod.rs:813-881): random pair sampled and folded in,
    derandomized, spartan::snark proof accepted by the restated verifier.  CPU twin: tests/test_r1cs_mirror_cpu.py."""
    from r1cs_parity import run_compressed_half
    run_compressed_half(b200, oracle, cid, num_cons, num_vars, 2, device_transcript)


@pytest.mark.parametrize("num_cons,num_vars,device_transcript", [(8, 8, False), (64, 32, True), (4, 16, True)])
def test_full_snark_with_hyperkzg_on_device(b200, oracle, num_cons, num_vars, device_transcript):
    """RelaxedR1CSSNARK::prove incl. EE::prove (HyperKZG with the transcript) on the device: proof equal to the
    oracle's and accepted by the restated verifier incl. the KZG opening equation.  CPU twin:
    tests/test_snark_mirror_cpu.py."""
    from snark_parity import run_full
    run_full(b200, oracle, num_cons, num_vars, 2, device_transcript)


@pytest.mark.parametrize("cid,num_cons,num_vars,device_transcript", [(1, 8, 8, False), (1, 32, 16, True), (3, 16, 16, True)])
def test_full_snark_with_ipa_on_device(b200, oracle, cid, num_cons, num_vars, device_transcript):
    """S2 of CompressedSNARK: spartan::snark + the IPA evaluation engine on the secondary curve (Grumpkin / Vesta),
    proof equal to the oracle's and accepted by the restated verifier.  CPU twin: tests/test_snark_mirror_cpu.py."""
    from snark_parity import run_full_ipa
    run_full_ipa(b200, oracle, cid, num_cons, num_vars, 2, device_transcript)


@pytest.mark.parametrize("num_cons,num_vars,device_transcript", [(8, 8, False), (16, 8, True), (64, 64, True)])
def test_whole_ppsnark_with_hyperkzg_on_device(b200, oracle, num_cons, num_vars, device_transcript):
    """ppsnark::RelaxedR1CSSNARK::prove incl. EE::prove (HyperKZG) over a test SRS; proof equal to the oracle's
    and accepted by the restated verifier incl. the batched opening.  CPU twin: tests/test_ppsnark_mirror_cpu.py."""
    import ppsnark_full_parity
    ppsnark_full_parity.run(b200, oracle, num_cons, num_vars, device_transcript)


def test_cpp_mirror_sumcheck_loops(oracle, tmp_path):
    """The C++ wrappers of the fused sum-check loops (prove_quad_prod, prove_cubic_with_three_inputs with a pending
    transcript buffer and a tau = 0 round) on the GPU; CPU twin in tests/test_cpp_mirror.py."""
    import test_cpp_mirror as tcm
    tcm.build()
    tcm.check_sumcheck(tcm.EXE, oracle, tmp_path)


def test_ptau_load_setup_on_device(b200, oracle, tmp_path):
    """PTAU file -> device-resident key (src/provider/ptau.rs, hyperkzg.rs:657-689): the reference's four read_ptau
    tests, header / truncation / non-canonical errors and the save_setup -> load_setup -> commit round trip, with
    the G1 section validated by k_on_curve in HBM (b200_ck_register_checked).  CPU twin: tests/test_ptau_cpu.py."""
    import ptau_parity
    from nova_b200 import ptau
    ptau_parity.run_reference_cases(ptau)
    ptau_parity.run_format_errors(ptau)
    ptau_parity.run_setup_round_trip(ptau, oracle, tmp_path)


def test_checked_registration_names_the_first_bad_point(b200, oracle):
    """b200_ck_register_checked on 2^14 points: clean key registers and commits; with two points corrupted the call
    fails with B200_E_POINT and first_bad = the smaller index; a bad blinding generator reports index n.
    CPU twin: tests/test_ptau_cpu.py."""
    import ptau_parity
    ptau_parity.run_checked_registration(oracle, 1 << 14, 9000, 12000)


def test_sharded_ptau_loading_two_ranks_on_device(tmp_path):
    """load_setup_sharded + sharded_commit with two gloo ranks driving the same GPU: slice keys validated in HBM,
    the combined commitment, the error agreement.  CPU twin: tests/test_ptau_sharded.py."""
    import test_ptau_sharded
    test_ptau_sharded.run_world(2, "gpu", tmp_path)


@pytest.mark.parametrize("cid", [1, 2, 3])
def test_pedersen_key_file_on_device(b200, oracle, cid):
    """Pedersen key file (pedersen.rs:317-340, 383-393) -> resident key on Grumpkin / Pallas / Vesta, validated in
    HBM; CPU twin: tests/test_ptau_cpu.py."""
    import ptau_parity
    from nova_b200 import ptau
    ptau_parity.run_pedersen_key_file(ptau, oracle, cid)


def test_sharded_hyperkzg_two_ranks_on_device(tmp_path):
    """sharded_hyperkzg_prove with two gloo ranks driving the same GPU (host-staged collectives): every prover
    message and the transcript equal to the unsharded oracle's.  CPU twin: tests/test_hyperkzg_sharded.py."""
    import test_hyperkzg_sharded
    test_hyperkzg_sharded.run_world(2, "gpu", tmp_path)


def test_sharded_hyperkzg_nccl(tmp_path):
    """the same over NCCL with one GPU per rank (NcclComm); needs at least two GPUs"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import test_hyperkzg_sharded
    test_hyperkzg_sharded.run_world(2, "nccl", tmp_path)


@pytest.mark.parametrize("seg", ["0", "1"])
def test_segmented_eq_reduction_on_device(seg):
    """NOVA_B200_SC_SEG=1 (k_form_reduce_eqseg: per-segment sums, pairs of indices sharing one reduction) gives
    the same t(0) / t(inf) / t(-1) sums as the default kernel and the oracle, on every eq-weighted form at a size
    where it is taken; in a subprocess because the switch is read once.  CPU twin: tests/test_host_kernels.py
    (the kernel itself on host threads)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, NOVA_B200_SC_SEG=seg)
    out = subprocess.run([sys.executable, os.path.join(here, "sc_seg_worker.py")], env=env, capture_output=True,
                         text=True, timeout=280)
    assert out.returncode == 0 and ("SEG OK" if seg == "1" else "FLAT OK") in out.stdout, out.stdout + out.stderr


def test_sharded_batched_sumcheck_two_ranks_on_device(tmp_path):
    """ppsnark.prove_helper_sharded with two gloo ranks driving the same GPU: cyclic shards, b200_sc_eval_sharded_dev
    for all nine sums, one exchange per round, replicated tail.  CPU twin: tests/test_ppsnark_sharded.py."""
    import test_ppsnark_sharded
    test_ppsnark_sharded.run_world(2, "gpu", tmp_path)


@pytest.mark.parametrize("fid", [0, 3])
def test_multi_evaluate_with_on_device(sp, fid):
    """b200_mle_eval_multi_dev (multi_evaluate_with, multilinear.rs:129-180) with the reference's test cases incl.
    the known values; CPU twin: tests/test_spartan_mirror_cpu.py."""
    import mle_multi_parity
    mle_multi_parity.run(sp, fid)


def _run_peer_world(world, kind, tmp_path):
    import subprocess
    import sys as _sys
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29700 + (os.getpid() % 1500) + world * 13
    out = str(tmp_path / f"peer_{world}")
    procs = [subprocess.Popen([_sys.executable, os.path.join(here, "peer_msm_worker.py"), str(r), str(world), str(port),
                               kind, out]) for r in range(world)]
    try:
        for pr in procs:
            assert pr.wait(timeout=300) == 0
    finally:
        for pr in procs:
            if pr.poll() is None:
                pr.kill()
    for r in range(world):
        assert open(f"{out}.{r}").read() == "OK"


@pytest.mark.parametrize("world", [2, 3])
def test_fused_sharded_msm_ranks_on_one_device(world, tmp_path):
    """b200_msm_sharded_dev: the reduction's last kernel publishes the rank's partial into every peer's exchange
    buffer (CUDA IPC), waits and sums -- here between processes that share cuda:0; results must equal the closed
    form of the whole vector and be bit-identical on every rank, over several epochs and with an empty rank."""
    _run_peer_world(world, "gpu", tmp_path)


def test_fused_sharded_msm_nccl_ranks(tmp_path):
    """the same with one GPU per rank (peer stores over NVLink); needs >= 2 GPUs"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_peer_world(2, "nccl", tmp_path)


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_one_process_multi_gpu_commit(devices, tmp_path):
    """b200_mgpu_*: one process, one call, the key block-cyclic over `devices` (here virtual devices that share
    cuda:0 -- the same code path as N physical GPUs: per-device streams, keys, workspaces, peer exchange inside the
    reduction kernels).  Run in a subprocess: the multi-GPU layer is initialised once per process."""
    import subprocess
    import sys as _sys
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([_sys.executable, os.path.join(here, "mgpu_commit_worker.py"), ",".join(map(str, devices))],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_one_process_all_gpus_commit(tmp_path):
    """the same over every physical GPU of the box (needs >= 2)"""
    import subprocess
    import sys as _sys

    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([_sys.executable, os.path.join(here, "mgpu_commit_worker.py"), ",".join(map(str, range(n)))],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
@pytest.mark.parametrize("arity", [5, 24])
def test_poseidon_ro_on_device(b200, fid, arity):
    """k_poseidon_ro (csrc/poseidon.cuh) == oracle/poseidon_ref.py: all four fields, the wide and the narrow sponge,
    inputs around the rate (several permutations while absorbing), two consecutive squeezes."""
    import poseidon_parity
    from nova_b200 import poseidon
    poseidon.PoseidonConstants._cache.clear()
    poseidon_parity.run_ro(b200, fid, arity)


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_poseidon_nifs_challenge_on_device(b200, oracle, cid):
    import poseidon_parity
    poseidon_parity.run_nifs_challenge(b200, oracle, cid)


def test_sharded_provers_over_nccl(tmp_path):
    """SURVEY §8e rows beyond the MSM on REAL multi-GPU hardware (one GPU per rank, NCCL): the batched inner sum-check of
    ppsnark on cyclic shards (prove_helper_sharded), the outer cubic sum-check, and the SpMV / cross-term / folding-step /
    Horner / division pieces -- each equal to the unsharded oracle message for message.  Needs >= 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import test_ppsnark_sharded
    import test_sharding_pieces
    import test_sumcheck_sharded
    test_ppsnark_sharded.run_world(2, "nccl", tmp_path)
    test_sharding_pieces.run_world(2, "nccl", tmp_path)
    test_sumcheck_sharded.run_world(2, "nccl", 0, 10, 0, tmp_path)


@pytest.mark.parametrize("fid", [0, 2])
def test_poly_eval_many(sp, oracle, fid):
    """b200_poly_eval_many_dev: polynomials on both sides of the short / long split (2^12 coefficients), the empty one, more
    than 32 short ones (two launches of k_poly_eval_small_multi), each at three points, against the oracle's Horner values."""
    import ctypes
    from nova_b200.native import check, lib
    lens = [0, 1, 2, 255, 256, 257, 4095, 4096, 4097, 70001] + [3 + i for i in range(34)]
    polys = [oracle.gen_scalars(fid, 900 + i, n) if n else b"" for i, n in enumerate(lens)]
    us = oracle.gen_scalars(fid, 33, 3)
    dev = [sp.DeviceVec.from_bytes(f) if f else None for f in polys]
    k = len(lens)
    ptrs = (ctypes.c_void_p * k)(*[d.ptr.value if d else None for d in dev])
    ud, ev = sp.DeviceVec.from_bytes(us), sp.DeviceVec(96 * k)
    check(lib().b200_poly_eval_many_dev(fid, ptrs, (ctypes.c_size_t * k)(*lens), k, ud.ptr, 3, ev.ptr, None))
    got = ev.to_bytes(96 * k)
    for i, f in enumerate(polys):
        want = oracle.poly_eval(fid, f, us) if f else bytes(96)
        assert got[96 * i:96 * i + 96] == want, lens[i]


def test_batched_sumcheck_concurrent_callers(b200, oracle):
    """Four host threads run b200_sumcheck_batched (each with its own tables, eq instances and transcript) while a fifth
    commits on a shared key: the calls share the library stream, the auxiliary stream (inversions) and the memory pool.
    Every proof must equal the oracle's prove_helper for its inputs."""
    import threading
    from nova_b200 import ppsnark as dp
    from nova_b200 import spartan as sp
    from oracle import ppsnark_ref as pr
    from oracle.pyref import CURVES
    fid, ell = 0, 9
    p, N = FIELD_MODULUS[fid], 1 << ell
    pack = lambda xs: b"".join(mont_bytes(p, x) for x in xs)

    def inputs(seed):
        rng = SplitMix64(seed)
        vec = lambda: [rng.field(p) for _ in range(N)]
        d = dict(oracles=[vec() for _ in range(4)], aux=[vec() for _ in range(4)], ts_row=vec(), ts_col=vec(), L_row=vec(),
                 L_col=vec(), val=vec(), E=vec(), W=vec(), rhos=[rng.field(p) for _ in range(ell)],
                 r_outer=[0 if i == 3 else rng.field(p) for i in range(ell)], claim=rng.field(p), claim_E=rng.field(p))
        return d

    def ref_run(d):
        mem = pr.MemorySumcheckInstance(p, d["oracles"], d["aux"], d["rhos"], d["ts_row"], d["ts_col"])
        inner = pr.InnerBatchedSumcheckInstance(p, d["claim"], d["L_row"], d["L_col"], d["val"], d["claim_E"], d["r_outer"], d["E"])
        wit = pr.WitnessBoundSumcheck(p, d["r_outer"], d["W"], 4)
        tr = Keccak256Transcript(p, b"cc")
        return pr.prove_helper(p, mem, inner, wit, tr), tr.squeeze(b"after")

    def dev_run(d):
        up = lambda v: sp.DeviceVec.from_bytes(pack(v))
        mem = dp.MemorySumcheckInstance(fid, N, [up(v) for v in d["oracles"]], [up(v) for v in d["aux"]], d["rhos"],
                                        up(d["ts_row"]), up(d["ts_col"]))
        inner = dp.InnerBatchedSumcheckInstance(fid, N, d["claim"], up(d["L_row"]), up(d["L_col"]), up(d["val"]), d["claim_E"],
                                                d["r_outer"], up(d["E"]))
        wit = dp.WitnessBoundSumcheck(fid, N, d["r_outer"], up(d["W"]), 4)
        tr = Keccak256Transcript(p, b"cc")
        return dp.prove_helper_device(fid, mem, inner, wit, tr), tr.squeeze(b"after")

    cases = [inputs(9000 + i) for i in range(4)]
    want = [ref_run(d) for d in cases]
    cid = 0
    n_ck = 1 << 12
    bases = oracle.gen_bases(cid, n_ck)
    ck = b200.CommitmentKey(b200.Curve(cid), bases)
    sc = oracle.gen_scalars(CURVES[cid].scalar_field, 77, n_ck)
    from test_msm_gpu import aff
    grp = b200.DlogGroup(cid)
    want_c = aff(CURVES[cid], oracle.msm(cid, sc, bases))
    got, errs, commits = [None] * 4, [], []

    def worker(i):
        try:
            for _ in range(3):
                got[i] = dev_run(cases[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def committer():
        try:
            for _ in range(12):
                commits.append(grp.vartime_multiscalar_mul(sc, ck))
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    ths = [threading.Thread(target=worker, args=(i,)) for i in range(4)] + [threading.Thread(target=committer)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for i in range(4):
        (g, after), (w, wafter) = got[i], want[i]
        assert [list(q) for q in g[0]] == [list(q) for q in w[0]] and list(g[1]) == list(w[1]) and g[2:] == w[2:], i
        assert after == wafter
    assert all(c == want_c for c in commits)
    ck.release()
