#!/usr/bin/env bash
# ncu captures of the dominant kernel (k_accumulate) for the default library and, if built, the fused-y3 variant.
# One GPU, run on the GPU box:   gpurun --timeout 900 -- 'bash tools/ncu_accumulate.sh'
# Outputs under gpurun_out/ncu/: .ncu-rep files (read them back here with `ncu -i ... --page raw --csv`) and a
# CSV of the headline metrics.  Numbers printed by a run under ncu are never bench values.
set -u
OUT=gpurun_out/ncu
mkdir -p "$OUT"
METRICS=sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
capture() {  # name, library
  local name=$1 libpath=$2
  NOVA_B200_LIB=$libpath ncu --set full --clock-control none --import-source on -k regex:k_accumulate -c 2 \
      -o "$OUT/accumulate_$name" -f python tools/devtime.py 20 > "$OUT/accumulate_$name.log" 2>&1
  NOVA_B200_LIB=$libpath ncu --metrics "$METRICS" --clock-control none -k regex:k_accumulate -c 2 --csv \
      --log-file "$OUT/accumulate_${name}_metrics.csv" python tools/devtime.py 20 > /dev/null 2>&1
}
capture base nova_b200/libnova_b200.so
[ -f nova_b200/libnova_b200_y3.so ] && capture y3 nova_b200/libnova_b200_y3.so
ls -la "$OUT"
