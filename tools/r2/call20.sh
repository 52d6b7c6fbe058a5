#!/bin/bash
timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py tests/test_ppsnark_gpu.py tests/test_abi.py -q -x -p no:cacheprovider -k "batched or ppsnark or cubic3 or snark or abi or export or concurrent" 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
