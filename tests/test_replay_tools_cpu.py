"""The prover replay tools (tools/ppsnark_replay.py -- also called by bench.py for its `snark_replays` side
measurement -- and tools/snark_replay.py) at toy size on the emulated device: a smoke test of their host
code paths (setup, synthetic shapes, both transcript modes), so that a typo cannot surface only on the GPU."""
import gc
import os
import sys

import pytest

import emulated_device

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def emulated():
    import nova_b200
    emulated_device.install()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, ROOT)
    yield nova_b200
    sys.path.remove(os.path.join(ROOT, "tools"))
    gc.collect()
    emulated_device.uninstall()


@pytest.mark.parametrize("device_transcript", [False, True])
def test_ppsnark_replay_runs(emulated, oracle, device_transcript):
    import ppsnark_replay
    out = ppsnark_replay.run(log2cons=4, reps=1, device_transcript=device_transcript)
    assert out["N"] >= 16 and out["ms"]["total"] > 0


@pytest.mark.parametrize("device_transcript", [False, True])
def test_snark_replay_runs(emulated, oracle, device_transcript):
    from tools import snark_replay
    out = snark_replay.run(log2cons=4, reps=1, device_transcript=device_transcript)
    assert out["ms"]["total"] > 0 and out["ms"]["hyperkzg_prove"] > 0
