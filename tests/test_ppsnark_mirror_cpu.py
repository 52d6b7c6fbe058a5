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


@pytest.mark.parametrize("cid,num_cons,num_vars", [(0, 8, 8), (1, 16, 8), (3, 4, 16)])
def test_ppsnark_prove_core_host_logic(emulated, oracle, cid, num_cons, num_vars):
    import test_ppsnark_gpu
    test_ppsnark_gpu.test_prove_core_matches_oracle(emulated, oracle, cid, num_cons, num_vars)
