#!/usr/bin/env bash
# full GPU suite + smoke + default bench line + launch list and ncu capture of the final state
set -u
OUT=gpurun_out/r2full
rm -rf "$OUT"; mkdir -p "$OUT/ncu"
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) | tee "$OUT/gpu_suite.txt"
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee "$OUT/smoke.txt"
( timeout 900 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc $?"; cut -c1-600 "$OUT/bench_n1.json" ) | tee -a "$OUT/gpu_suite.txt"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/ncu/launches_bench_2p20.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prove-step --other-log2n "" > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_accumulate -s 2 -c 1 -o /tmp/ncu_acc -f python tools/devtime.py 20 > "$OUT/ncu/acc_full.log" 2>&1
ncu -i /tmp/ncu_acc.ncu-rep --page raw --csv > "$OUT/ncu/accumulate_full_raw.csv" 2>/dev/null
du -sh "$OUT"
