import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from oracle import model as O
from sensor_probe import make_ctx
sensors = sys.argv[1]
m = O.Model("biped-4x3", "flat", sensors, O.CTRL_MLP_CENTRALIZED, O.default_params(1.0, 0.2))
g = np.random.default_rng(12).uniform(-1, 1, m.genome_len())
_, otv, otb, otc = m.run(g, trace_steps=20)
for prec in (64, 32):
    ctx = make_ctx(m, prec)
    _, tv, tb, tc = ctx.trace(g, 20)
    ctx.close()
    for k in range(20):
        eb = np.abs(tb[k] - otb[k])
        i = np.unravel_index(np.argmax(eb), eb.shape)
        print(prec, "step", k + 1, "ctrl err %.2e" % np.abs(tc[k] - otc[k]).max(), "argmax voxel", int(np.argmax(np.abs(tc[k] - otc[k]))),
              "body err %.2e at body %d comp %d" % (eb.max(), i[0], i[1]), "ctrl", np.round(tc[k][:4], 4), np.round(otc[k][:4], 4))
