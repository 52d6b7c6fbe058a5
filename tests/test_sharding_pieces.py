"""The remaining multi-GPU pieces of SURVEY.md §8e -- SpMV / folding-step cross term on row slices (one
all-gather of z), Horner evaluation and division by (X - u) on index-range slices (one all-gather of a field
element per rank) -- over gloo with the oracle as the per-rank engine (CPU) and with the device engine (GPU).
Every rank's slice must equal the corresponding slice of the unsharded result."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run_world(world, kind, tmp_path):
    port = 29900 + (os.getpid() % 1500) + world * 11
    out = str(tmp_path / f"pieces_{world}")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "shard_pieces_worker.py"), str(r), str(world), str(port),
                               kind, out]) for r in range(world)]
    try:
        for pr in procs:
            assert pr.wait(timeout=300) == 0
    finally:
        for pr in procs:  # a failed or stuck rank must not leave its peers waiting in a collective
            if pr.poll() is None:
                pr.kill()
    for r in range(world):
        assert open(f"{out}.{r}").read() == "OK"


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pieces_gloo_cpu(world, tmp_path):
    run_world(world, "oracle", tmp_path)


def test_sharded_pieces_device_engine_on_emulated_device(tmp_path):
    """nova_b200.sharding.DeviceEngine (the adapters a GPU rank uses) with the library emulated on the CPU."""
    run_world(2, "emulated", tmp_path)

