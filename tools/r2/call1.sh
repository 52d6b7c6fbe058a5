#!/usr/bin/env bash
# round-2 call 1: measure what round 1 left unmeasured (A/B variants, H2D chunks, segmented eq, replays) + ncu of the tail kernels
set -u
OUT=gpurun_out/r2c1
rm -rf "$OUT"; mkdir -p "$OUT"
run() { local name=$1; shift; echo "== $name: $*" | tee -a "$OUT/summary.txt"; ( "$@" ) > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; grep -E "stages|msm best|\"ms\"|passed|failed|error|pct|GB/s" "$OUT/$name.log" | tail -12 | tee -a "$OUT/summary.txt"; }
nvidia-smi > "$OUT/smi.txt" 2>&1
run base_devtime   timeout 300  python tools/devtime.py 17 20 22
run y3_devtime     timeout 300  env NOVA_B200_LIB=nova_b200/libnova_b200_y3.so python tools/devtime.py 17 20 22
run y3sq_devtime   timeout 300  env NOVA_B200_LIB=nova_b200/libnova_b200_y3sq.so python tools/devtime.py 17 20 22
run y3_parity      timeout 600  env NOVA_B200_LIB=nova_b200/libnova_b200_y3.so python -m pytest tests/test_msm_gpu.py -m gpu -x -q -p no:cacheprovider
run y3sq_parity    timeout 600  env NOVA_B200_LIB=nova_b200/libnova_b200_y3sq.so python -m pytest tests/test_msm_gpu.py tests/test_fieldvec_gpu.py -m gpu -x -q -p no:cacheprovider
for c in 13 14 15 16; do run dev17_c$c timeout 120 python tools/devtime.py 17 c=$c; done
run e2e_base       timeout 200  python tools/e2e_commit.py --log-n 20
for k in 2 4 8; do run e2e_chunks$k timeout 200 env NOVA_B200_H2D_CHUNKS=$k python tools/e2e_commit.py --log-n 20; done
run e2e22_base     timeout 200  python tools/e2e_commit.py --log-n 22 --reps 8
run e2e22_chunks4  timeout 200  env NOVA_B200_H2D_CHUNKS=4 python tools/e2e_commit.py --log-n 22 --reps 8
run scflat_sumcheck timeout 300 python tools/sumcheck_replay.py --log-n 22 --reps 3
run scseg_sumcheck  timeout 300 env NOVA_B200_SC_SEG=1 python tools/sumcheck_replay.py --log-n 22 --reps 3
run ppsnark_host   timeout 600  python tools/ppsnark_replay.py --log2cons 18 --reps 2
run ppsnark_dev    timeout 600  python tools/ppsnark_replay.py --log2cons 18 --reps 2 --device-transcript
run ppsnark_seg    timeout 600  env NOVA_B200_SC_SEG=1 python tools/ppsnark_replay.py --log2cons 18 --reps 2
run fieldbench     timeout 300  python tools/fieldbench.py
# ncu: full-set captures of the tail kernels (1 launch each), converted to CSV on the box (the .ncu-rep stays there)
mkdir -p "$OUT/ncu"
for k in k_scatter k_red_final_q k_red_digits_q k_fixup_q k_digits; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -o /tmp/ncu_$k -f python tools/devtime.py 20 > "$OUT/ncu/$k.log" 2>&1
  ncu -i /tmp/ncu_$k.ncu-rep --page raw --csv > "$OUT/ncu/${k}_raw.csv" 2>/dev/null
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/ncu/launches_devtime20.csv" python tools/devtime.py 20 > /dev/null 2>&1
du -sh "$OUT"
