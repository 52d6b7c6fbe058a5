"""Host logic of nova_b200/r1cs.py (commit_T / commit_T_relaxed, NIFS orchestration, witness and instance
folds, is_sat_relaxed: r1cs/mod.rs:474-664, 1044-1107, 1237-1292; nifs.rs:36-167) on the CPU through
tests/emulated_device.py, on the reference's own folding fixture (nifs.rs:299-351)."""
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


@pytest.mark.parametrize("cid", [0, 1, 2, 3])
def test_fold_twice_then_relaxed_sat_host_logic(emulated, oracle, cid):
    from r1cs_parity import run_tiny_fixture
    run_tiny_fixture(emulated, oracle, cid)


@pytest.mark.parametrize("cid", [0, 2])
def test_streamed_witness_folding_steps_host_logic(emulated, oracle, cid):
    from r1cs_parity import run_streamed_steps
    run_streamed_steps(emulated, oracle, cid)


@pytest.mark.parametrize("cid,device_transcript", [(0, False), (0, True), (1, False)])
def test_compressed_snark_half_host_logic(emulated, oracle, cid, device_transcript):
    from r1cs_parity import run_compressed_half
    run_compressed_half(emulated, oracle, cid, device_transcript=device_transcript)
