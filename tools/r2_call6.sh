#!/bin/bash
# round 2, GPU call 6: per-voxel sensor lists / lidar tests, all GPU tests with the shipped arithmetic (libm sincosf / tanhf, refresh),
# fresh measured bounds, bench with other_configs
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
rm -f $O/parity_measured.jsonl
python -m pytest tests -m gpu -q --tb=short > $O/c6_pytest.log 2>&1; echo "pytest rc $?" >> $O/c6_pytest.log
tail -30 $O/c6_pytest.log | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 > $O/c6_bench.json 2> $O/c6_bench.err; echo "bench rc $?"; tail -3 $O/c6_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c6_bench.json').read().strip().splitlines()[-1])
print('value %.4e e2e %.4e ms %.2f roofline %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))
print('ga', d['generation_loop']['ms_per_generation'], d['generation_loop']['voxel_steps_per_sec'])
for o in d.get('other_configs', []): print(o['workload'][:30], o['robots'], '%.4e' % o['voxel_steps_per_sec'], o['kernel_ms'], o['kernel'])
print('parity', d.get('parity'))
PY
python tools/evolved_probe.py time final
