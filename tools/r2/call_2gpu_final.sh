#!/usr/bin/env bash
# final tree on two physical GPUs: the four multi-GPU twins + the sharded bench lines
set -u
OUT=gpurun_out/r2final2
rm -rf "$OUT"; mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_zz_new_paths_gpu.py -q -p no:cacheprovider -m gpu -k "nccl or one_process or fused_sharded" 2>&1 | tail -5 ) | tee "$OUT/tests_2gpu.txt"
( timeout 300 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-1500 ) | tee "$OUT/bench_n2.json"
( timeout 300 $TR --master-port 29522 bench.py --workload hyperkzg --log2n 22 --steps 3 --warmup 2 --gpus 2 2>&1 | tail -1 | cut -c1-1500 ) | tee "$OUT/hyperkzg_2p22_n2.json"
du -sh "$OUT"
