#!/bin/bash
# usage (GPU box): tools/ab_cfg.sh <workload> <robots> name1 name2 ... -> bench of every build_ab/<name> ("base" = the in-tree library), two rounds
wl=$1; robots=$2; shift 2
for round in 1 2; do
  for n in "$@"; do
    if [ "$n" = base ]; then e="X=1"; else e="VSR_LIB=build_ab/$n/libvsr_eval.so"; fi
    echo -n "workload $wl robots $robots build $n: "
    env $e python bench.py --workload $wl --robots $robots --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs --ga-generations 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print("value %.3e  ms/step %7.1f" % (d["value"], d["ms_per_step"]))
"
  done
done
