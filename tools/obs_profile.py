"""One vsr_eval_ex launch with the average-posture request (config 2, 3552 robots = one wave of the 8-warp observation build): ncu target."""
import sys
import numpy as np
sys.path.insert(0, ".")
from vsrevolution_b200.locomotion import population_evaluator
ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 30.0, 1.0)
g = np.random.default_rng(0).uniform(-1, 1, (3552, ev.genome_len))
for _ in range(2):
    ev.ctx.eval_ex(g, posture_n=8)
print(ev.ctx.last_kernel_ms(), ev.ctx.kernel_info())
