#!/usr/bin/env bash
# round-2 call 4 (1 GPU): slice-pipelined e2e, L_MIN sweep for small shards, peer exchange between processes on one device,
# prover timings with the pool + native Keccak, launch list of one HyperKZG proof
set -u
OUT=gpurun_out/r2c4
rm -rf "$OUT"; mkdir -p "$OUT/ncu"
run() { local name=$1; shift; echo "== $name: $*" | tee -a "$OUT/summary.txt"; ( "$@" ) > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; grep -E "stages|msm best|passed|failed|rror|\"metric\"|\"what\"" "$OUT/$name.log" | cut -c1-1200 | tail -8 | tee -a "$OUT/summary.txt"; }
run peer_tests     timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -p no:cacheprovider -k "fused_sharded"
run msm_tests      timeout 900 python -m pytest tests/test_msm_gpu.py tests/test_fieldvec_gpu.py -m gpu -q -p no:cacheprovider
for k in 1 2 3 4; do run e2e20_slices$k timeout 200 env NOVA_B200_E2E_SLICES=$k python tools/e2e_commit.py --log-n 20; done
for k in 1 4; do run e2e22_slices$k timeout 200 env NOVA_B200_E2E_SLICES=$k python tools/e2e_commit.py --log-n 22 --reps 8; done
for k in 1 4; do run e2e24_slices$k timeout 300 env NOVA_B200_E2E_SLICES=$k python tools/e2e_commit.py --log-n 24 --reps 4; done
for l in 32 24 16 12 8; do run dev17_lmin$l timeout 120 env NOVA_B200_ACC_LMIN=$l python tools/devtime.py 17 18; done
run wl_hkzg22      timeout 600 python bench.py --workload hyperkzg --log2n 22 --steps 3
run wl_ppsnark18   timeout 600 python bench.py --workload ppsnark --log2cons 18 --steps 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file "$OUT/ncu/launches_hkzg20.csv" python tools/workloads.py hyperkzg --log2 20 --steps 1 > "$OUT/ncu/hkzg20.log" 2>&1
python - <<'PY' | tee -a gpurun_out/r2c4/summary.txt
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2c4/ncu/launches_hkzg20.csv")) if len(r)>14 and r[0].isdigit()]
tot=collections.Counter(); cnt=collections.Counter()
for r in rows:
    name=r[4].split("<")[0].split("(")[0].replace("void ","").replace("nova::","")
    tot[name]+=float(r[14])/1e3; cnt[name]+=1
print("launch list of the HyperKZG 2^20 workload run (all proofs incl. warm-up + check): kernel, launches, total us")
for k,v in tot.most_common(25): print(f"  {k:40s} {cnt[k]:6d} {v:12.1f}")
PY
du -sh "$OUT"
