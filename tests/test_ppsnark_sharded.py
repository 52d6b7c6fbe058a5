"""The batched inner sum-check of MicroSpartan over several ranks (nova_b200.ppsnark.prove_helper_sharded;
BASELINE.json configs[4], SURVEY.md §8e "Sum-check round"): sixteen tables sharded cyclically, nine partial sums
exchanged per round, binds local, replicated tail; every prover message, the final claims and the transcript
state equal the unsharded oracle's, incl. rounds whose tau is 0.  gloo, one emulated device per rank (CPU); GPU
variant in tests/test_zz_new_paths_gpu.py."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run_world(world, kind, tmp_path):
    port = 23500 + (os.getpid() % 1500) + world * 19
    out = str(tmp_path / f"pps_{world}")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "shard_ppsnark_worker.py"), str(r), str(world),
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
def test_sharded_batched_sumcheck_gloo_cpu(world, tmp_path):
    run_world(world, "emulated", tmp_path)
