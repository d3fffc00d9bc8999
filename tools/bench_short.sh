#!/bin/bash
# usage: tools/bench_short.sh "<env assignments>" -> prints value and kernel info
env $1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-other-configs 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    k=d['kernel']
    print('%-40s value %.3e  ms/step %6.1f  e2e %.3e  regs %d ctas/sm %d thr %d smem %d' % ('$1', d['value'], d['ms_per_step'], d['e2e']['value'], k['regs_per_thread'], k['ctas_per_sm'], k['threads_per_cta'], k['smem_bytes']))
"
