"""nova_b200.poseidon (constants derivation, RO bookkeeping, resident squeeze plumbing) on the emulated device: the
library calls are answered by oracle/poseidon_ref.py, which also REJECTS a registration whose constants differ from
its own -- so the product's derivation is compared element by element on the way in.  CPU only; the kernel itself is
covered by the GPU twin in tests/test_zz_new_paths_gpu.py."""
import gc

import pytest

import emulated_device
import poseidon_parity


@pytest.fixture()
def emulated():
    import nova_b200
    from nova_b200 import poseidon
    emulated_device.install()
    poseidon.PoseidonConstants._cache.clear()
    yield nova_b200
    poseidon.PoseidonConstants._cache.clear()
    gc.collect()
    emulated_device.uninstall()


@pytest.mark.parametrize("fid,arity", [(0, 24), (1, 5), (3, 24)])
def test_ro_host_logic(emulated, fid, arity):
    poseidon_parity.run_ro(emulated, fid, arity)


@pytest.mark.parametrize("cid", [0, 1, 2])
def test_nifs_challenge_host_logic(emulated, oracle, cid):
    poseidon_parity.run_nifs_challenge(emulated, oracle, cid)
