#!/bin/bash
# concurrency test of the batched loop + ncu --set full of its two large-round kernels (round 0 at 2^20 entries)
mkdir -p gpurun_out/r02l
timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -x -p no:cacheprovider -m gpu -k "concurrent_callers or batched" 2>&1 | tail -3
for k in k_form_reduce_multi k_bind_top_multi; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -c 1 -o /tmp/ncu_$k -f python tools/r2/batched_once.py 20 0 > gpurun_out/r02l/$k.log 2>&1
  ncu -i /tmp/ncu_$k.ncu-rep --page raw --csv > gpurun_out/r02l/r02l_ncu_${k}_raw.csv 2>/dev/null
  python - "$k" <<'PY'
import csv, sys
k = sys.argv[1]
rows = list(csv.reader(open(f'gpurun_out/r02l/r02l_ncu_{k}_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.sum', 'smsp__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'lts__t_sector_hit_rate.pct', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
print('==', k)
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f'  {w:75s} {vals[i]:>16s} {units[i]}')
PY
done
du -sh gpurun_out/r02l
