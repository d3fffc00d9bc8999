"""BASELINE config 5: the biped-4x3 MLP evaluator inside the numGA generation loop, population 2^12..2^18, on the GPUs of this
launch (python tools/ga_scaling.py for 1 GPU, torchrun --nproc-per-node N for N).  One JSON line per population size."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from bench import WORKLOADS
from vsrevolution_b200.locomotion import population_evaluator
from vsrevolution_b200.solver import num_ga

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = WORKLOADS[2]
ev = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], w["T"], 1.0, devices=[local])
pops = [int(a) for a in sys.argv[1:]] or [1 << e for e in range(12, 19)]
for n_pop in pops:
    loop = num_ga().build({"nPop": str(n_pop), "nEval": "0"}).build(ev, seed=1, process_group=dist.group.WORLD if world > 1 else None)
    loop.init()
    loop.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    gens = 2
    l0 = loop.ga.kernel_launches
    t0 = time.perf_counter()
    for _ in range(gens):
        loop.step()
    best = loop.ga.best_fitness()
    dt = time.perf_counter() - t0
    kms = ev.ctx.last_kernel_ms()
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        per = dt / gens
        print(json.dumps({"n_gpus": world, "n_pop": n_pop, "s_per_generation": per, "evals_per_s": n_pop / per,
                          "voxel_steps_per_s": n_pop * ev.n_voxels * 1801 / per, "episode_kernel_ms_rank0": kms,
                          "loop_overhead_frac": max(0.0, 1 - kms * 1e-3 / per), "best_fitness": best,
                          "ga_kernel_launches_per_generation": (loop.ga.kernel_launches - l0) / gens}))
        sys.stdout.flush()
    loop.ga.close()
if world > 1:
    dist.destroy_process_group()
