#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1 -f -o $O/r2b_prof_cfg3 python bench.py --workload 3 --robots 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 > $O/f_ncu_3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 3 -c 1 -f -o $O/r2b_prof_cfg4 python bench.py --workload 4 --robots 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 > $O/f_ncu_4.log 2>&1
ls -la $O/r2b_*.ncu-rep
