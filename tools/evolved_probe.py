"""Episode-kernel time on an EVOLVED config-2 population (evolved genomes simulate ~15 % slower than random ones).
  python tools/evolved_probe.py make [gens]   -> gpurun_out/evolved_g.npy (4096 robots after `gens` numGA generations, T = 30 s)
  python tools/evolved_probe.py time <label>  -> kernel ms on the random and on the evolved genomes for the library / launch shape /
                                                 iteration counts selected by the environment (VSR_LIB, VSR_LAUNCH_SHAPE, VSR_POS_ITERS...)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from vsrevolution_b200 import _ffi
from vsrevolution_b200.locomotion import population_evaluator

cmd = sys.argv[1]
T = 30.0
path = "build_ab/evolved_g.npy"  # git-ignored, travels to the GPU box (gpurun_out/ does not)
if cmd == "make":
    gens = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", T, 1.0)
    ga = _ffi.GeneticAlgorithm(ev.ctx, 4096, seed=0)
    ga.init()
    for _ in range(gens):
        ga.step()
    g, f = ga.population()
    os.makedirs("build_ab", exist_ok=True)
    np.save(path, g)
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/evolved_g.npy", g)
    print(json.dumps({"made": path, "gens": gens, "best": float(f[0]), "median": float(np.median(f)), "kernel_ms_last_gen": ev.ctx.last_kernel_ms()}))
else:
    label = sys.argv[2] if len(sys.argv) > 2 else ""
    e = _ffi.env_default()
    e.final_t, e.transient_t = T, 1.0
    if os.environ.get("VSR_POS_ITERS"):
        e.pos_iters = int(os.environ["VSR_POS_ITERS"])
    if os.environ.get("VSR_VEL_ITERS"):
        e.vel_iters = int(os.environ["VSR_VEL_ITERS"])
    xs, ys = _ffi.terrain("flat")
    ctx = _ffi.Context(e, xs, ys)
    ctx.body_set(_ffi.shape_mask("biped-4x3"), _ffi.sensor_config("uniform-a-0"))
    ctx.controller_set(_ffi.CTRL_MLP_CENTRALIZED, [10, 7, 10])
    rec = {"label": label}
    sets = {"random": np.random.default_rng(0).uniform(-1, 1, (4096, 157))}
    if os.path.exists(path):
        sets["evolved"] = np.load(path)
    for k, g in sets.items():
        ms = []
        for _ in range(4):
            ctx.eval(g)
            ms.append(ctx.last_kernel_ms())
        rec[k + "_ms"] = round(float(np.min(ms[1:])), 3)
        import hashlib
        rec[k + "_fit_sha"] = hashlib.sha1(ctx.eval(g).tobytes()).hexdigest()[:12]  # bit-identity across builds
    rec["info"] = ctx.kernel_info()
    ctx.close()
    print(json.dumps(rec))
