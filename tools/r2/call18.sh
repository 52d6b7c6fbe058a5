#!/bin/bash
timeout 600 python tools/prove_step_replay.py ab 2>&1 | tail -4
timeout 300 python bench.py --workload prove_step --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('prove_step', d['value'], d['parity_checked'])"
