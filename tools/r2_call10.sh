#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python tools/obs_overhead.py > $O/obs_overhead.json 2>&1; cat $O/obs_overhead.json | head -40
python tools/ga_scaling.py > $O/r2_ga_scaling_1gpu.jsonl 2> $O/ga1.err; cut -c1-200 $O/r2_ga_scaling_1gpu.jsonl
