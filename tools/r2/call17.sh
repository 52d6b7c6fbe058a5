#!/bin/bash
for occ in 2 3; do
  echo "== NOVA_B200_SC_MULTI_OCC=$occ"
  NOVA_B200_SC_MULTI_OCC=$occ timeout 300 python tools/r2/batched_time.py 20 2>&1 | tail -3
done
NOVA_B200_SC_MULTI_OCC=3 timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -x -p no:cacheprovider -m gpu -k "batched or ppsnark" 2>&1 | tail -2
for occ in 2 3; do
  NOVA_B200_SC_MULTI_OCC=$occ timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); ph=d['detail']['phases_ms']; print('ppsnark occ', '$occ', d['value'], 'inner', ph['inner_sumcheck'], 'outer', ph['outer_sumcheck'], d['parity_checked'])"
done
