#!/bin/bash
echo "== round kernel, called multiplier"; ./tools/roundbench_call
echo "== round kernel, inlined multiplier"; ./tools/roundbench_inline
timeout 600 python tools/r2/batched_time.py 12 16 2>&1 | tail -16
timeout 900 python -m pytest tests/test_zz_new_paths_gpu.py tests/test_ppsnark_gpu.py -q -x -p no:cacheprovider -m gpu -k "ppsnark or batched or cubic3 or quad_prod or snark" 2>&1 | tail -4
