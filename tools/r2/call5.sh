#!/bin/bash
# bind_multi + C++ mgpu wrapper verification, then ppsnark timing
mkdir -p gpurun_out
timeout 1100 python -m pytest tests/test_ppsnark_gpu.py tests/test_zz_new_paths_gpu.py tests/test_cpp_mirror.py tests/test_sumcheck_gpu.py -q -x -p no:cacheprovider -m gpu 2>&1 | tail -8
for i in 1 2; do timeout 300 python bench.py --workload ppsnark --log2cons 18 --steps 3 --warmup 3 2>&1 | tail -1; done
