#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
./tools/microbench/ffma2_latency > $O/r2_microbench_ffma2_latency.txt 2>&1; cat $O/r2_microbench_ffma2_latency.txt
VARIANTS="head f2s f2w" bash tools/r2_call12.sh
