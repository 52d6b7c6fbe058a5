#!/bin/bash
timeout 1200 python -m pytest tests/test_zz_new_paths_gpu.py tests/test_ppsnark_gpu.py tests/test_spartan_gpu.py -q -x -p no:cacheprovider -m gpu -k "poly_eval or hyperkzg or ppsnark or snark or batched" 2>&1 | tail -4
timeout 300 python bench.py --workload hyperkzg --log2n 22 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('hyperkzg', d['value'], json.dumps(d['detail'].get('phases_ms')), d['parity_checked'])"
timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ppsnark', d['value'], d.get('gpu_launches'), json.dumps(d['detail']['phases_ms']), d['parity_checked'])"
