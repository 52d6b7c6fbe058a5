"""PTAU loading for one process per GPU (nova_b200.ptau.load_setup_sharded): every rank
reads, uploads and validates only its index range of the TauG1 section; commitments through the slice keys equal
the unsharded commitment; validation errors are agreed on by all ranks.  gloo, emulated device per rank (CPU);
GPU variant in tests/test_zz_new_paths_gpu.py."""
import io
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def make_file(tmp_path):
    """a 128-point test SRS with tau_H, written by the product's write_ptau"""
    sys.path.insert(0, os.path.dirname(HERE))
    from nova_b200 import ptau
    from oracle import hyperkzg_ref as hk
    srs = hk.setup_srs(0, 128, 0xC0FFEE)
    tau_H = ptau.g2_to_raw(ptau.g2_mul(ptau.G2_GENERATOR, 0xC0FFEE))
    buf = io.BytesIO()
    ptau.write_ptau(buf, srs, tau_H + tau_H, 8)
    path = tmp_path / "srs128.ptau"
    path.write_bytes(buf.getvalue())
    return str(path)


def run_world(world, kind, tmp_path):
    path = make_file(tmp_path)
    port = 27100 + (os.getpid() % 1500) + world * 13
    out = str(tmp_path / f"ptau_{world}")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "shard_ptau_worker.py"), str(r), str(world),
                               str(port), kind, path, out]) for r in range(world)]
    try:
        for pr in procs:
            assert pr.wait(timeout=300) == 0
    finally:
        for pr in procs:
            if pr.poll() is None:
                pr.kill()
    for r in range(world):
        assert open(f"{out}.{r}").read() == "OK"


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ptau_loading_gloo_cpu(world, tmp_path):
    run_world(world, "emulated", tmp_path)
