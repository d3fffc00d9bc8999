"""Kernel time of sub-wave batches of config 2 (latency regime: the 255-register build, few warps per SM) for the library in VSR_LIB."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from vsrevolution_b200.locomotion import population_evaluator
ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 30.0, 1.0)
rec = {"label": sys.argv[1] if len(sys.argv) > 1 else ""}
import hashlib
for n in (3, 96, 512, 2048, 3552):
    g = np.random.default_rng(0).uniform(-1, 1, (n, ev.genome_len))
    ms = []
    for _ in range(3):
        fit = ev.ctx.eval(g); ms.append(ev.ctx.last_kernel_ms())
    rec[f"n{n}_ms"] = round(min(ms[1:]), 2)
    rec[f"n{n}_sha"] = hashlib.sha1(fit.tobytes()).hexdigest()[:8]
print(json.dumps(rec))
