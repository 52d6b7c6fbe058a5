#!/usr/bin/env bash
# round-2 scaling run (gpurun --gpus 8): the sharded MSM at 1/2/4/8 GPUs (fused peer exchange), HyperKZG 2^22 at 1/2/4/8 (NCCL,
# verified proofs), one-process multi-GPU commit over all devices, peer exchange between processes on separate GPUs
set -u
OUT=gpurun_out/r2s8
rm -rf "$OUT"; mkdir -p "$OUT"
run() { local name=$1; shift; echo "== $name: $*" | tee -a "$OUT/summary.txt"; ( "$@" ) > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; grep -E "passed|failed|rror|\"metric\"" "$OUT/$name.log" | cut -c1-400 | tail -4 | tee -a "$OUT/summary.txt"; }
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run msm_n1 timeout 300 python bench.py --no-cpu-baseline --no-prove-step
for n in 2 4 8; do run msm_n$n timeout 300 $TR --nproc-per-node $n --master-port $((29520+n)) bench.py --gpus $n; done
run msm_n8_nccl timeout 300 $TR --nproc-per-node 8 --master-port 29539 bench.py --gpus 8 --exchange nccl --other-log2n ""
run hkzg_n1 timeout 400 python bench.py --workload hyperkzg --log2n 22 --steps 3
for n in 2 4 8; do run hkzg_n$n timeout 400 $TR --nproc-per-node $n --master-port $((29540+n)) bench.py --workload hyperkzg --log2n 22 --steps 3 --gpus $n; done
run mgpu_tests timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -p no:cacheprovider -k "one_process_all or fused_sharded_msm_nccl"
python - <<'PY' | tee -a gpurun_out/r2s8/summary.txt
import json
def last_json(f):
    try:
        for line in open(f):
            if line.startswith("{"): d=json.loads(line)
        return d
    except Exception as e:
        return None
print("MSM 2^20 (device ms | e2e ms | stages | other sizes | check)")
base=None
for n in (1,2,4,8):
    d=last_json(f"gpurun_out/r2s8/msm_n{n}.log")
    if not d: print(n,"missing"); continue
    if n==1: base=d["ms_per_step"]
    print(n, round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), "speedup", round(base/d["ms_per_step"],2) if base else None, d["roofline"]["stage_ms_per_msm"], [(o["log2n"],o["ms_per_step"]) for o in d["other_sizes"]], d["result_check"]["ok"])
d=last_json("gpurun_out/r2s8/msm_n8_nccl.log")
if d: print("8 nccl", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4))
print("HyperKZG 2^22")
for n in (1,2,4,8):
    d=last_json(f"gpurun_out/r2s8/hkzg_n{n}.log")
    if d: print(n, d["value"], d["e2e"]["value"], d["parity_checked"], d["detail"]["parity"])
PY
