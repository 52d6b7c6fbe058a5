"""Host logic of the provider mirror's remaining CommitmentEngine methods on the CPU (emulated device)."""
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


@pytest.mark.parametrize("cid", [0, 2])
def test_commit_variants_host_logic(emulated, oracle, cid):
    import commit_variants_parity
    commit_variants_parity.run(emulated, oracle, cid)
