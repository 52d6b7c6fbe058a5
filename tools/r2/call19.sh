#!/bin/bash
# launch list of one verified ppsnark proof (2^18 constraints): which kernels the 61 ms are made of
mkdir -p gpurun_out/r02m
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02m/r02m_launches_ppsnark_2p18.csv python tools/r2/ppsnark_once.py > gpurun_out/r02m/run.log 2>&1
tail -2 gpurun_out/r02m/run.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02m/r02m_launches_ppsnark_2p18.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
seq = []
for r in rows[1:]:
    name = r[ki].split('<')[0].split('(')[0].replace('void ', '')
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    seq.append((name, v))
# the proof of interest = launches after the LAST k_spmv burst start (the script runs setup, one warm proof, one measured proof)
marks = [i for i, (n, _) in enumerate(seq) if n == 'k_logup_hash']
start = marks[len(marks) // 2] if marks else 0
# walk back to the spmv launches that open the proof
while start > 0 and seq[start - 1][0] != 'k_spmv': start -= 1
while start > 0 and seq[start - 1][0] == 'k_spmv': start -= 1
part = seq[start:]
agg = collections.OrderedDict()
for n, v in part:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(t for _, t in agg.values())
print(f'launches in the last proof: {len(part)}, kernel time {tot/1e3:.2f} ms')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f'  {k:34s} {n:5d} launches {t/1e3:8.3f} ms {100*t/tot:5.1f} %')
PY
