#!/usr/bin/env bash
# One gpurun call that settles everything round 1 left unmeasured (its GPU budget ran out after the last
# kernels were written).  Run from the repo root on the GPU box:
#     gpurun --timeout 1500 -- 'bash tools/gpu_checklist.sh'
# Everything lands in gpurun_out/checklist/ ; nothing here is a bench value (bench.py is the bench).
# Before calling: build the A/B variant HERE (it travels as a .so):
#     make -C nova_b200/csrc variant VARIANT=y3 VFLAGS=-DNOVA_MADD_FUSED_Y3 -j8
#     make -C nova_b200/csrc variant VARIANT=y3sq "VFLAGS=-DNOVA_MADD_FUSED_Y3 -DNOVA_SQR_DEDICATED" -j8
set -u
OUT=gpurun_out/checklist
mkdir -p "$OUT"
run() { local name=$1; shift; echo "== $name: $*" | tee -a "$OUT/summary.txt"; ( "$@" ) > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; }

# 1. parity: the tests that have never seen a GPU (last-sorted file), then the whole GPU suite
run zz_new_paths     timeout 600  python -m pytest tests/test_zz_new_paths_gpu.py -q -rA --timeout 300 -p no:cacheprovider
run gpu_suite        timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider

# 2. fused-y3 mixed addition: parity under the variant library, then stage timings A/B
if [ -f nova_b200/libnova_b200_y3.so ]; then
  run y3_parity      timeout 900  env NOVA_B200_LIB=nova_b200/libnova_b200_y3.so python -m pytest tests/test_msm_gpu.py -m gpu -x -q -p no:cacheprovider
  run y3_devtime     timeout 300  env NOVA_B200_LIB=nova_b200/libnova_b200_y3.so python tools/devtime.py 20 22
  run base_devtime   timeout 300  python tools/devtime.py 20 22
else
  echo "libnova_b200_y3.so missing: build the variant first" | tee -a "$OUT/summary.txt"
fi

if [ -f nova_b200/libnova_b200_y3sq.so ]; then   # + dedicated squaring (static profile: fewer IMAD, more registers)
  run y3sq_parity    timeout 900  env NOVA_B200_LIB=nova_b200/libnova_b200_y3sq.so python -m pytest tests/test_msm_gpu.py tests/test_fieldvec_gpu.py -m gpu -x -q -p no:cacheprovider
  run y3sq_devtime   timeout 300  env NOVA_B200_LIB=nova_b200/libnova_b200_y3sq.so python tools/devtime.py 20 22
fi

# 3. prover replays: host transcript vs device transcript, streamed witness reuse
run sumcheck_replay  timeout 300  python tools/sumcheck_replay.py --log-n 20 --reps 3
run snark_dev        timeout 600  python tools/snark_replay.py --log2cons 20 --reps 2
run snark_host       timeout 600  python tools/snark_replay.py --log2cons 20 --reps 2 --host-transcript
run ppsnark_replay   timeout 600  python tools/ppsnark_replay.py --log2cons 18 --reps 2
run sumcheckeq       timeout 600  python tools/sumcheckeq_replay.py --min 10 --max 22 --reps 2
run ppsnark_dev      timeout 600  python tools/ppsnark_replay.py --log2cons 18 --reps 2 --device-transcript

# 3b. segmented eq reductions (NOVA_B200_SC_SEG=1): parity of the sum-check suites under the switch, then A/B timing
run scseg_parity     timeout 900  env NOVA_B200_SC_SEG=1 python -m pytest tests/test_spartan_gpu.py tests/test_ppsnark_gpu.py -m gpu -x -q -p no:cacheprovider
run scseg_sumcheck   timeout 300  env NOVA_B200_SC_SEG=1 python tools/sumcheck_replay.py --log-n 22 --reps 3
run scflat_sumcheck  timeout 300  python tools/sumcheck_replay.py --log-n 22 --reps 3
run scseg_ppsnark    timeout 600  env NOVA_B200_SC_SEG=1 python tools/ppsnark_replay.py --log2cons 18 --reps 2

# 4. end-to-end commit: chunked upload overlapping the digit stage (tuning hook, off by default)
run e2e_base         timeout 300  python tools/e2e_commit.py --log-n 20
run e2e_chunks4      timeout 300  env NOVA_B200_H2D_CHUNKS=4 python tools/e2e_commit.py --log-n 20
run e2e_chunks8      timeout 300  env NOVA_B200_H2D_CHUNKS=8 python tools/e2e_commit.py --log-n 20
run e2e_chunks4_par  timeout 900  env NOVA_B200_H2D_CHUNKS=4 python -m pytest tests/test_msm_gpu.py -m gpu -x -q -p no:cacheprovider

# 5. multi-GPU box only (gpurun --gpus 2): HyperKZG over two GPUs, NCCL and host-staged collectives
NG=$(python -c "import torch; print(torch.cuda.device_count())" 2>/dev/null || echo 1)
if [ "$NG" -ge 2 ]; then
  run hkzg_1gpu      timeout 600  python tools/hyperkzg_sharded_replay.py --log2n 22
  run hkzg_2gpu_nccl timeout 600  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/hyperkzg_sharded_replay.py --log2n 22 --comm nccl
  run hkzg_2gpu_host timeout 600  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/hyperkzg_sharded_replay.py --log2n 22 --comm host
fi

grep -h "passed\|failed\|error" "$OUT"/zz_new_paths.log "$OUT"/gpu_suite.log "$OUT"/y3_parity.log 2>/dev/null | tail -6 | tee -a "$OUT/summary.txt"
