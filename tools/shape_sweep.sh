#!/bin/bash
# usage (GPU box): tools/shape_sweep.sh -> launch-shape A/B on configs 3 and 4 (VSR_LAUNCH_SHAPE: 0 = 4 warps x 2 CTAs at 255 regs,
# 1 = 5 x 2 at 168, 2 = 8 x 1 at 255, 3 = 10 x 1 at 168; "auto" = the library's own choice)
for wl in 3 4; do
  for robots in 2960 16384; do
    for shape in auto 0 1 2 3; do
      if [ "$shape" = auto ]; then e="X=1"; else e="VSR_LAUNCH_SHAPE=$shape"; fi
      echo -n "workload $wl robots $robots shape $shape: "
      env $e python bench.py --workload $wl --robots $robots --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=d['kernel']
    print('value %.3e  ms/step %7.1f  regs %d ctas/sm %d thr %d smem %d' % (d['value'], d['ms_per_step'], k['regs_per_thread'], k['ctas_per_sm'], k['threads_per_cta'], k['smem_bytes']))
"
    done
  done
done
