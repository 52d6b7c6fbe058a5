#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/r2/batched_time.py 12 16 20 2>&1 | tail -30
for tb in 0 8; do
  echo "== ppsnark NOVA_B200_SC_TAIL_BITS=$tb"
  NOVA_B200_SC_TAIL_BITS=$tb timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d.get('gpu_launches'), json.dumps(d['detail']['phases_ms']), d['detail']['N'])"
done
