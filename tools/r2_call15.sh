#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python tools/obs_profile.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vsr_episode -s 1 -c 1 -f -o $O/r2_prof_obs python tools/obs_profile.py > $O/ncu_obs.log 2>&1; tail -3 $O/ncu_obs.log
ls -la $O/r2_prof_obs.ncu-rep
