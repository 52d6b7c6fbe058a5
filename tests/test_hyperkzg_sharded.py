"""HyperKZG EvaluationEngine::prove over several ranks (nova_b200.sharding.sharded_hyperkzg_prove; SURVEY.md §8d
config C4 "sharded 1/2/4/8 GPUs", §8e): polynomial split by index range, every prover message (fold commitments,
3-point evaluations, quotient commitments) and the transcript state equal to the unsharded oracle's.  gloo, one
emulated device per rank (CPU); GPU variant in tests/test_zz_new_paths_gpu.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run_world(world, kind, tmp_path):
    port = 25300 + (os.getpid() % 1500) + world * 17
    out = str(tmp_path / f"hkzg_{world}")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "shard_hyperkzg_worker.py"), str(r), str(world),
                               str(port), kind, out]) for r in range(world)]
    try:
        for pr in procs:
            assert pr.wait(timeout=300) == 0
    finally:
        for pr in procs:
            if pr.poll() is None:
                pr.kill()
    for r in range(world):
        assert open(f"{out}.{r}").read() == "OK"


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_hyperkzg_gloo_cpu(world, tmp_path):
    run_world(world, "emulated", tmp_path)
