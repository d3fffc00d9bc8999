#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q --tb=short -k "multi_device" > $O/g2_pytest.log 2>&1; tail -3 $O/g2_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/g2_bench.json 2> $O/g2_bench.err; echo "bench rc $?"; tail -3 $O/g2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g2_bench.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.4e e2e %.4e ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print('ga', d['generation_loop'])
print('multi', d.get('multi_device_exact'))
print('parity', {k: d['parity'][k] for k in ('spearman','median_abs_dfitness','leading_steps_within_1e-4_of_voxel_side')})
PY
