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


@pytest.mark.parametrize("num_vars", [3, 5])
def test_sumcheckeq_replay_matches_oracle(emulated, oracle, num_vars):
    """The reference's `sumcheckeq` bench workload (benches/sumcheckeq.rs:24-115) through the mirror equals the
    oracle's MemorySumcheckInstance round by round (taus[0] = 0 exercises the fall-back)."""
    import sumcheckeq_replay
    from oracle import ppsnark_ref as pr
    from oracle.pyref import FIELD_MODULUS
    p = FIELD_MODULUS[0]
    n = 1 << num_vars
    _, got = sumcheckeq_replay.run_sc(num_vars, collect=True)
    vs = [[i * k % p for i in range(n)] for k in range(1, 9)]
    inst = pr.MemorySumcheckInstance(p, vs[0:4], vs[4:8], [(-2 * i) % p for i in range(num_vars)], vs[0], vs[1])
    for j in range(num_vars):
        assert [list(e) for e in got[j]] == [list(e) for e in inst.evaluation_points()], j
        inst.bound((-j) % p)


def test_sharded_hyperkzg_replay_prints_the_single_gpu_digest(emulated, oracle):
    """tools/hyperkzg_sharded_replay.py as a single rank and tools/hyperkzg_replay.py agree on the digest of the
    proof (same key, polynomial and challenges) -- the cross-check the two tools are meant to give on the GPU."""
    import hyperkzg_replay
    import hyperkzg_sharded_replay
    a = hyperkzg_sharded_replay.main(["--log2n", "6", "--reps", "1", "--comm", "host"])
    b = hyperkzg_replay.gpu(log2n=6, reps=1)
    assert a["digest"] == b["digest"] and a["n_gpus"] == 1


def test_hyperkzg_workload_is_checked_by_the_verifier(emulated, oracle):
    """tools/workloads.hyperkzg (bench.py --workload hyperkzg) at toy size on the emulated device: the proof is
    accepted by the restated verifier for C, y computed by the oracle, a tampered one is rejected."""
    import workloads
    out = workloads.hyperkzg(log2n=5, steps=1, warmup=0)
    assert out["parity_checked"] and out["parity"]["verifier_accepts"] and out["parity"]["verifier_rejects_tampered"]
    assert out["ms_per_proof"] > 0 and out["e2e_ms_per_proof"] > 0


@pytest.mark.parametrize("device_transcript", [False, True])
def test_ppsnark_workload_is_checked_by_the_verifier(emulated, oracle, device_transcript):
    import workloads
    out = workloads.ppsnark(log2cons=4, steps=1, warmup=0, device_transcript=device_transcript)
    assert out["parity_checked"], out["parity"]
    assert out["parity"]["instance_commitments_equal_oracle"] and out["parity"]["verifier_rejects_tampered"]
