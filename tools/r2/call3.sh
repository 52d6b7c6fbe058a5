#!/usr/bin/env bash
# round-2 call 3 (gpurun --gpus 2): NCCL / peer-memory paths on two real GPUs + pooled-allocator timings
set -u
OUT=gpurun_out/r2c3
rm -rf "$OUT"; mkdir -p "$OUT"
run() { local name=$1; shift; echo "== $name: $*" | tee -a "$OUT/summary.txt"; ( "$@" ) > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; grep -E "stages|msm best|passed|failed|rror|\"metric\"" "$OUT/$name.log" | cut -c1-1800 | tail -8 | tee -a "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
run peer_tests     timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -x -p no:cacheprovider -k "fused_sharded or sharded_hyperkzg or sharded_pieces or sharded_batched or sharded_ptau"
run wl_hkzg22_n1   timeout 600 python bench.py --workload hyperkzg --log2n 22 --steps 3
run wl_hkzg22_n2   timeout 600 $TR --master-port 29511 bench.py --workload hyperkzg --log2n 22 --steps 3 --gpus 2
run wl_hkzg20_n2   timeout 600 $TR --master-port 29512 bench.py --workload hyperkzg --log2n 20 --steps 3 --gpus 2
run wl_ppsnark18   timeout 600 python bench.py --workload ppsnark --log2cons 18 --steps 3
run msm_n1         timeout 600 python bench.py --no-cpu-baseline --no-prove-step
run msm_n2_fused   timeout 600 $TR --master-port 29513 bench.py --gpus 2
run msm_n2_nccl    timeout 600 $TR --master-port 29514 bench.py --gpus 2 --exchange nccl
du -sh "$OUT"
