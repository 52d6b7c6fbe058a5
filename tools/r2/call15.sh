#!/bin/bash
timeout 600 python -m pytest tests/test_spartan_gpu.py tests/test_ppsnark_gpu.py -q -x -p no:cacheprovider -m gpu -k "invert or ppsnark or oracles" 2>&1 | tail -3
for ch in 32 64 128; do
  echo "== NOVA_B200_BINV_CHUNK=$ch"
  NOVA_B200_BINV_CHUNK=$ch timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); ph=d['detail']['phases_ms']; print('ppsnark', d['value'], 'memory_oracles', ph['memory_oracles'], 'final_evals_rlc', ph['final_evals_rlc'], 'total', ph['total'], d['parity_checked'])"
done
