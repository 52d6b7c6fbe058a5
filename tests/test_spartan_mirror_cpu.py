"""Host logic of nova_b200/spartan.py on the CPU through tests/emulated_device.py: the bodies of the GPU
parity tests for the sum-check round loops (incl. tau = 0 fall-backs and prove_batch_eval with instances
of different sizes) and for the HyperKZG prover core run unchanged against the emulated device."""
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


def test_sumcheck_round_loops_host_logic(emulated, oracle):
    import test_spartan_gpu as t
    from nova_b200 import spartan
    t.test_sumcheck_provers_match_reference_restatement(spartan, oracle, 0)
    t.test_prove_batch_eval_matches_restatement(spartan, oracle)


def test_hyperkzg_prove_core_host_logic(emulated, oracle):
    import test_spartan_gpu as t
    from nova_b200 import spartan
    t.test_hyperkzg_prove_core(emulated, spartan, oracle)


@pytest.mark.parametrize("k,l,zero", [(1, 3, ()), (3, 5, ()), (4, 6, (0, 3)), (16, 2, (1,))])
def test_prove_batched_cubic_host_logic(emulated, oracle, k, l, zero):
    import batched_cubic_parity
    from nova_b200 import spartan
    batched_cubic_parity.run(spartan, 0, k, l, zero)


@pytest.mark.parametrize("cid,l", [(1, 1), (1, 5), (3, 4), (0, 3)])
def test_ipa_prover_host_logic(emulated, oracle, cid, l):
    """nova_b200/ipa.py (the prover that never folds the key) on the emulated device: body of tests/test_ipa_gpu.py."""
    import test_ipa_gpu
    test_ipa_gpu.test_ipa_prove_matches_restatement(emulated, oracle, cid, l)


@pytest.mark.parametrize("fid", [0, 3])
def test_multi_evaluate_with_host_logic(emulated, oracle, fid):
    import mle_multi_parity
    from nova_b200 import spartan
    mle_multi_parity.run(spartan, fid)
