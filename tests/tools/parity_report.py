"""Parity evidence for DESIGN.md section 6 (run on the GPU box): (1) the number of leading steps N for which the CUDA voxel
vertices stay within 1e-4 (relative to the voxel side) of the oracle, FP64 and FP32 builds, configs 1/3/4; (2) fitness agreement
over the full config-2 population (4096 robots x 30 s): Spearman rank correlation, median |d|, fractions within tolerances."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import model as O
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity_probe import make_ctx

SIDE = 3.0
out = {"tolerance": "max over voxel vertices of |cuda - oracle| <= 1e-4 * voxel side (3.0)", "first_steps": {}, "population": {}}
for cfg, seed in ((1, 0), (3, 1), (4, 2)):
    m = O.config_model(cfg, final_t=6.0)
    g = np.random.default_rng(seed).uniform(-1, 1, (8, m.genome_len()))
    steps = 300
    rec = {}
    for prec in (64, 32):
        ctx = make_ctx(m, prec)
        ns = []
        for r in range(8):
            _, otv, _, _ = m.run(g[r], trace_steps=steps)
            _, tv, _, _ = ctx.trace(g[r], steps)
            err = np.abs(tv - otv).reshape(steps, -1).max(axis=1) / SIDE
            bad = np.nonzero(err > 1e-4)[0]
            ns.append(int(bad[0]) if len(bad) else steps)
        rec["fp%d" % prec] = {"N_per_robot": ns, "N_min": min(ns), "N_median": float(np.median(ns))}
        ctx.close()
    out["first_steps"]["config%d" % cfg] = rec
    print(cfg, rec); sys.stdout.flush()

def spearman(a, b):
    ra, rb = np.argsort(np.argsort(a)), np.argsort(np.argsort(b))
    return float(np.corrcoef(ra, rb)[0, 1])

m = O.config_model(2, final_t=30.0)
n = 4096
g = np.random.default_rng(0).uniform(-1, 1, (n, m.genome_len()))
t0 = time.perf_counter()
import os
ref = m.eval_batch(g, n_threads=os.cpu_count() or 8)
out["population"]["oracle_seconds"] = time.perf_counter() - t0
fits = {}
for prec in (64, 32):
    ctx = make_ctx(m, prec)
    fits[prec] = ctx.eval(g)
    ctx.close()
def cmp(a, b):
    d = np.abs(a - b)
    return {"spearman": spearman(a, b), "median_abs_diff": float(np.median(d)), "mean_abs_diff": float(d.mean()),
            "frac_within_1e-3": float((d < 1e-3).mean()), "frac_within_1e-2": float((d < 1e-2).mean()),
            "mean_a": float(a.mean()), "mean_b": float(b.mean()), "std_a": float(a.std()), "std_b": float(b.std())}
out["population"].update({"robots": n, "episode_s": 30.0, "units": "fitness = velocity (voxel sides are 3 units; velocities are units/s)",
                          "fp32_vs_oracle": cmp(fits[32], ref), "fp64_vs_oracle": cmp(fits[64], ref), "fp32_vs_fp64": cmp(fits[32], fits[64])})
print(json.dumps(out["population"], indent=1))
json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1)
