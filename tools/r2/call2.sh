#!/usr/bin/env bash
# round-2 call 2: TMA-staged accumulate A/B (parity + stage times + ncu), the verifier-checked workloads at scale, full bench line
set -u
OUT=gpurun_out/r2c2
rm -rf "$OUT"; mkdir -p "$OUT/ncu"
run() { local name=$1; shift; echo "== $name: $*" | tee -a "$OUT/summary.txt"; ( "$@" ) > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; grep -E "stages|msm best|\"ms\"|passed|failed|rror|parity_checked|\"metric\"" "$OUT/$name.log" | cut -c1-1500 | tail -12 | tee -a "$OUT/summary.txt"; }
run base_devtime   timeout 300  python tools/devtime.py 17 20 22
run tma_devtime    timeout 300  env NOVA_B200_ACC_TMA=1 python tools/devtime.py 17 20 22
run tma_parity     timeout 600  env NOVA_B200_ACC_TMA=1 python -m pytest tests/test_msm_gpu.py -m gpu -x -q -p no:cacheprovider
run wl_prove_step  timeout 300  python bench.py --workload prove_step --steps 5
run wl_hyperkzg20  timeout 400  python bench.py --workload hyperkzg --log2n 20 --steps 3
run wl_hyperkzg22  timeout 600  python bench.py --workload hyperkzg --log2n 22 --steps 3
run wl_ppsnark16   timeout 400  python bench.py --workload ppsnark --log2cons 16 --steps 2
run wl_ppsnark18   timeout 600  python bench.py --workload ppsnark --log2cons 18 --steps 2
run bench_default  timeout 900  python bench.py
METRICS=sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_uniform.sum
timeout 300 ncu --metrics "$METRICS" --clock-control none -k regex:k_accumulate -s 2 -c 1 --csv --log-file "$OUT/ncu/acc_base.csv" python tools/devtime.py 20 > /dev/null 2>&1
timeout 300 env NOVA_B200_ACC_TMA=1 ncu --metrics "$METRICS" --clock-control none -k regex:k_accumulate -s 2 -c 1 --csv --log-file "$OUT/ncu/acc_tma.csv" python tools/devtime.py 20 > /dev/null 2>&1
for v in base tma; do
  if [ $v = tma ]; then export NOVA_B200_ACC_TMA=1; fi
  timeout 300 ncu --set full --clock-control none -k regex:k_accumulate -s 2 -c 1 -o /tmp/ncu_acc_$v -f python tools/devtime.py 20 > "$OUT/ncu/acc_full_$v.log" 2>&1
  ncu -i /tmp/ncu_acc_$v.ncu-rep --page raw --csv > "$OUT/ncu/acc_full_${v}_raw.csv" 2>/dev/null
done
unset NOVA_B200_ACC_TMA
du -sh "$OUT"
