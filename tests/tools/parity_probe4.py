"""Dev probe (GPU box): per-step FP64-kernel vs oracle error on other terrains (first steps after landing)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import model as O
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity_probe import make_ctx
names = sys.argv[1:] or ["hilly-1-10-42", "hilly-1-10-5"]
g = np.random.default_rng(4).uniform(-1, 1, (8, 16))
idx = {"hilly-1-10-5": 4, "hilly-1-10-42": 5}
for name in names:
    m = O.config_model(3, final_t=2.0)
    xs, ys = O.create_terrain(name)
    m.tx, m.ty = np.ascontiguousarray(xs), np.ascontiguousarray(ys)
    ctx = make_ctx(m, 64)
    i = idx.get(name, 0)
    res, otv, otb, otc = m.run(g[i], trace_steps=60)
    _, tv, tb, tc = ctx.trace(g[i], 60)
    err = np.abs(tv - otv).reshape(60, -1).max(axis=1)
    berr = np.abs(tb - otb).reshape(60, -1, 6)
    print(name, "verts under robot:", [(round(x, 2), round(y, 2)) for x, y in zip(xs, ys) if 5 < x < 70])
    for k in range(20, 60):
        worst = np.unravel_index(np.argmax(berr[k]), berr[k].shape)
        print(f"  step {k+1}: vert err {err[k]:.2e}  worst body {worst[0]} (voxel {worst[0]//4} mass {worst[0]%4}) comp {worst[1]} err {berr[k][worst]:.2e}  manifolds/step so far oracle={res.n_manifolds}")
    ctx.close()
