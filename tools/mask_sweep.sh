#!/bin/bash
# usage (GPU box): tools/mask_sweep.sh <workload> <robots> mask1 mask2 ... -> phase-barrier mask A/B (VSR_SYNC_MASK, bits in vsr_kernel.cuh KParams)
wl=$1; robots=$2; shift 2
for m in "$@"; do
  echo -n "workload $wl robots $robots mask $m: "
  env VSR_SYNC_MASK=$m python bench.py --workload $wl --robots $robots --steps 2 --warmup 3 --no-cpu-baseline --ga-generations 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('value %.3e  ms/step %7.1f' % (d['value'], d['ms_per_step']))
"
done
