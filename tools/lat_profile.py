"""One config-2 launch of 96 robots (32 robot-warps, one per SM: the latency regime): ncu target."""
import sys
import numpy as np
sys.path.insert(0, ".")
from vsrevolution_b200.locomotion import population_evaluator
ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 30.0, 1.0)
g = np.random.default_rng(0).uniform(-1, 1, (96, ev.genome_len))
for _ in range(2):
    ev.ctx.eval(g)
print(ev.ctx.last_kernel_ms(), ev.ctx.kernel_info())
