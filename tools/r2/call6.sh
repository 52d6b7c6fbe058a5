#!/bin/bash
# native batched sum-check loop: parity, then ppsnark timing per tail setting
mkdir -p gpurun_out
timeout 1300 python -m pytest tests/test_ppsnark_gpu.py tests/test_zz_new_paths_gpu.py tests/test_cpp_mirror.py -q -x -p no:cacheprovider -m gpu -k "ppsnark or batched or cubic3 or quad_prod or snark or mirror" 2>&1 | tail -8
for tb in 0 6 8 10; do
  echo "== NOVA_B200_SC_TAIL_BITS=$tb"
  NOVA_B200_SC_TAIL_BITS=$tb timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d.get('gpu_launches'), json.dumps(d.get('phases_ms') or d.get('config')))"
done
