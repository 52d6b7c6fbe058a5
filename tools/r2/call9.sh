#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_launches_batched_sumcheck_2p16.csv python tools/r2/batched_once.py 16 0 > gpurun_out/r02h_batched_once.log 2>&1
tail -3 gpurun_out/r02h_batched_once.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02h_launches_batched_sumcheck_2p16.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
seq = []
for r in rows[1:]:
    name = r[ki].split('<')[0].split('(')[0]
    v = float(r[vi].replace(',', '')); 
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    seq.append((name, v))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print(f"{k:40s} {n:5d} launches {t:10.1f} us total {t/n:8.2f} us avg")
print("last 24 launches:")
for name, v in seq[-24:]: print(f"   {name:36s} {v:8.2f} us")
PY
