#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $O/g${N}_bench.json 2> $O/g${N}_bench.err; echo "bench rc $?"; tail -2 $O/g${N}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/g${N}_bench.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value %.4e e2e %.4e ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
g=d['generation_loop']; print('ga ms/gen %.2f voxel-steps/s %.4e kernel %.2f' % (g['ms_per_generation'], g['voxel_steps_per_sec'], g['episode_kernel_ms']))
print('multi', d.get('multi_device_exact'))
PY
