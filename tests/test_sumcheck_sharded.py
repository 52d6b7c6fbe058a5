"""Sum-check sharded over ranks (cyclic layout, per-round all-gather of 2-3 field elements,
SURVEY.md §8e): every rank must emit exactly the proof of the unsharded big-integer restatement of
sumcheck.rs:446-507.  CPU: world 2 and 4 over gloo with the oracle as the per-rank engine.
GPU: world 2, both ranks on cuda:0, device engine."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run_world(world, engine, fid, l, zero_tau, tmp_path):
    port = 29600 + (os.getpid() % 1500) + world * 7 + l
    out = str(tmp_path / f"res_{world}_{l}_{zero_tau}")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "sumcheck_shard_worker.py"), str(r), str(world),
                               str(port), engine, str(fid), str(l), str(zero_tau), out]) for r in range(world)]
    for pr in procs:
        assert pr.wait(timeout=300) == 0
    for r in range(world):
        assert open(f"{out}.{r}").read() == "OK"


@pytest.mark.parametrize("world,l,zero_tau", [(2, 6, 0), (4, 7, 0), (2, 5, 1), (2, 1, 0)])
def test_sharded_sumcheck_gloo_cpu(world, l, zero_tau, tmp_path):
    run_world(world, "oracle", 0, l, zero_tau, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("world,l,zero_tau", [(2, 10, 0), (4, 9, 1)])
def test_sharded_sumcheck_gpu(world, l, zero_tau, tmp_path):
    run_world(world, "gpu", 0, l, zero_tau, tmp_path)
