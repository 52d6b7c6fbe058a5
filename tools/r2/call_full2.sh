#!/usr/bin/env bash
# final tree: full GPU suite + smoke + default bench line + single-loop sum-check timings
set -u
OUT=gpurun_out/r2final
rm -rf "$OUT"; mkdir -p "$OUT"
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) | tee "$OUT/gpu_suite.txt"
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee "$OUT/smoke.txt"
( timeout 900 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench rc $?"; cut -c1-700 "$OUT/bench_n1.json" ) | tee -a "$OUT/gpu_suite.txt"
timeout 300 python tools/sumcheck_replay.py --log-n 18 --reps 4 2>&1 | tail -4 | cut -c1-700 | tee "$OUT/sumcheck_replay_2p18.txt"
timeout 300 python tools/sumcheck_replay.py --log-n 22 --reps 4 2>&1 | tail -4 | cut -c1-700 | tee "$OUT/sumcheck_replay_2p22.txt"
du -sh "$OUT"
