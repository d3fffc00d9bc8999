#!/bin/bash
# usage (GPU box): tools/latency_probe.sh -> sub-wave batches: the 10-warp kernel against the 255-register kernel launched with few warps per CTA
for robots in 3 96 512 1024 2048; do
  for cfg in "X=1" "VSR_LAUNCH_SHAPE=2" "VSR_LAUNCH_SHAPE=2 VSR_WARPS=4" "VSR_LAUNCH_SHAPE=2 VSR_WARPS=2" "VSR_LAUNCH_SHAPE=2 VSR_WARPS=1" "VSR_LAUNCH_SHAPE=0 VSR_WARPS=2" "VSR_LAUNCH_SHAPE=0 VSR_WARPS=1"; do
    echo -n "robots $robots [$cfg]: "
    env $cfg python bench.py --workload 2 --robots $robots --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=d['kernel']
    print('ms/step %7.2f  value %.3e  regs %d thr %d ctas/sm %d' % (d['ms_per_step'], d['value'], k['regs_per_thread'], k['threads_per_cta'], k['ctas_per_sm']))
"
  done
done
