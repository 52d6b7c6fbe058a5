#!/bin/bash
echo "== round kernel"; ./tools/roundbench_call
timeout 600 python tools/r2/batched_time.py 12 16 20 2>&1 | tail -24
timeout 900 python -m pytest tests/test_zz_new_paths_gpu.py tests/test_ppsnark_gpu.py -q -x -p no:cacheprovider -m gpu -k "ppsnark or batched or cubic3 or quad_prod or snark" 2>&1 | tail -4
for tb in 0 8; do
  echo "== ppsnark NOVA_B200_SC_TAIL_BITS=$tb"
  NOVA_B200_SC_TAIL_BITS=$tb timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d.get('gpu_launches'), json.dumps(d['detail']['phases_ms']), d['parity_checked'])"
done
