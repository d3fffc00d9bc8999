"""Dev probe (GPU box): per-step FP64-kernel vs oracle error for an arbitrary (shape, ctrl, sensors, terrain)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import model as O
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity_probe import make_ctx
shape, ctrl, sensors, terrain = sys.argv[1:5]
kind = {"phase": O.CTRL_PHASE_SIN, "central": O.CTRL_MLP_CENTRALIZED, "dist": O.CTRL_MLP_DISTRIBUTED_HOMO}[ctrl]
m = O.Model(shape, terrain, sensors, kind, O.default_params(1.5, 0.5))
g = np.random.default_rng(11).uniform(-1, 1, (70, m.genome_len()))
ctx = make_ctx(m, 64)
for r in range(2):
    _, otv, otb, otc = m.run(g[r], trace_steps=40)
    _, tv, tb, tc = ctx.trace(g[r], 40)
    for k in range(40):
        be = np.abs(tb[k] - otb[k]); w = np.unravel_index(np.argmax(be), be.shape)
        print(f"robot {r} step {k+1}: body err {be.max():.2e} (body {w[0]} comp {w[1]})  ctrl err {np.abs(tc[k]-otc[k]).max():.2e}  min y {otb[k][:,1].min():.3f}")
