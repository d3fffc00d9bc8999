#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "observation or broken or launch_shape" > $O/c14_pytest.log 2>&1; tail -12 $O/c14_pytest.log | cut -c1-250
for n in 3552 4096; do
python tools/obs_overhead.py $n > $O/obs_overhead_$n.json 2>&1
python - $n <<'PY'
import json,sys
d=json.load(open('gpurun_out/obs_overhead_%s.json'%sys.argv[1]))
for k,v in d.items(): print(sys.argv[1], k, v['kernel_ms'], v['wall_ms'], v['info']['regs_per_thread'], v['info']['threads_per_cta'])
PY
done
