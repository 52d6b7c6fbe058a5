"""Host logic of nova_b200/snark.py (the mirror of spartan::snark::RelaxedR1CSSNARK::prove, snark.rs:113-256)
on the CPU: the library is replaced by tests/emulated_device.py, which answers every `*_dev` call with the
C oracle on host memory, so the mirror's glue (buffer sizes and offsets, argument order, transcript labels,
batch_eval_reduce bookkeeping) is compared message by message with oracle/snark_ref.py.  The CUDA kernels
themselves are covered by the `-m gpu` tests."""
import gc

import pytest

import emulated_device


@pytest.fixture()
def emulated():
    import nova_b200
    dev = emulated_device.install()
    yield nova_b200
    gc.collect()
    emulated_device.uninstall()


@pytest.mark.parametrize("cid,num_cons,num_vars,num_io", [(0, 4, 4, 1), (0, 16, 8, 2), (1, 8, 16, 2), (3, 32, 32, 3)])
@pytest.mark.parametrize("device_transcript", [False, True])
def test_snark_prove_core_host_logic(emulated, oracle, cid, num_cons, num_vars, num_io, device_transcript):
    """device_transcript=True: the emulated device strings the HOST BUILD of the round kernel together as
    csrc/capi_sumcheck.inc does, so the mirror's marshalling of the transcript is covered too."""
    from snark_parity import run_case
    run_case(emulated, oracle, cid, num_cons, num_vars, num_io, device_transcript=device_transcript)


@pytest.mark.parametrize("device_transcript", [False, True])
def test_full_snark_with_hyperkzg_host_logic(emulated, oracle, device_transcript):
    from snark_parity import run_full
    run_full(emulated, oracle, 8, 8, 2, device_transcript)


@pytest.mark.parametrize("cid,device_transcript", [(1, False), (1, True), (3, False)])
def test_full_snark_with_ipa_host_logic(emulated, oracle, cid, device_transcript):
    from snark_parity import run_full_ipa
    run_full_ipa(emulated, oracle, cid, 8, 8, 2, device_transcript)
