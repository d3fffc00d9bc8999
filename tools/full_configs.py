"""Run the BASELINE configs at their full sizes once (dev/validation): throughput, finiteness, pool overflow."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from bench import WORKLOADS
from vsrevolution_b200.locomotion import population_evaluator
from vsrevolution_b200 import _ffi
out = {}
for wl, n in ((3, 65536), (4, 16384), (2, 4096), (2, 16384), (2, 65536), (2, 262144)):
    w = WORKLOADS[wl]
    ev = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], w["T"], 1.0)
    g = np.random.default_rng(w["seed"]).uniform(-1, 1, (n, ev.genome_len))
    ev.ctx.eval(g)  # untimed: allocates the context's device and page-locked staging buffers for this batch size
    t0 = time.perf_counter()
    fit, res = ev.ctx.eval(g, want_results=True)
    dt = time.perf_counter() - t0
    steps = int(res["n_steps"][0])
    vs = n * ev.n_voxels * steps / dt
    rec = dict(workload=w["name"], robots=n, steps=steps, seconds=dt, kernel_ms=ev.ctx.last_kernel_ms(), voxel_steps_per_s=vs,
               evals_per_s=n / dt, finite=bool(np.isfinite(fit).all()), overflow_robots=int((res["contact_overflow"] > 0).sum()),
               fitness_mean=float(fit.mean()), fitness_std=float(fit.std()), fitness_max=float(fit.max()), fitness_min=float(fit.min()))
    print(json.dumps(rec)); sys.stdout.flush()
