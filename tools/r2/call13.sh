#!/bin/bash
timeout 900 python -m pytest tests/test_zz_new_paths_gpu.py tests/test_ppsnark_gpu.py -q -x -p no:cacheprovider -m gpu -k "ppsnark or batched or cubic3 or quad_prod or snark or sumcheck" 2>&1 | tail -4
timeout 300 python tools/sumcheck_replay.py --log-n 18 --reps 4 2>&1 | tail -4 | cut -c1-700
timeout 300 python tools/sumcheck_replay.py --log-n 22 --reps 4 2>&1 | tail -4 | cut -c1-700
