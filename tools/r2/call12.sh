#!/bin/bash
timeout 900 python -m pytest tests/test_zz_new_paths_gpu.py tests/test_ppsnark_gpu.py tests/test_sharding_gpu.py -q -x -p no:cacheprovider -m gpu -k "ppsnark or batched or cubic3 or quad_prod or snark or sumcheck" 2>&1 | tail -4
timeout 300 python tools/sumcheck_replay.py --log2n 18 --reps 4 2>&1 | tail -1 | cut -c1-900
timeout 300 python tools/sumcheck_replay.py --log2n 22 --reps 4 2>&1 | tail -1 | cut -c1-900
timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d.get('gpu_launches'), json.dumps(d['detail']['phases_ms']), d['parity_checked'])"
