"""Two ppsnark proofs at 2^18 constraints (one warm, one for the launch list), verified: for ncu --metrics gpu__time_duration"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import workloads as wl  # noqa: E402

res = wl.ppsnark(log2cons=18, steps=1, warmup=1)
print(json.dumps({"ms_per_proof_under_ncu": res["ms_per_proof"], "parity_checked": res["parity_checked"]}))
