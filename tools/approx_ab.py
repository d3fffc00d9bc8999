"""FP32-approximation A/B: for the library named by VSR_LIB, the number of leading steps the FP32 kernel's
voxel vertices stay within 1e-4 (relative to the voxel side) of the oracle's, 8 seeded robots on configs 1/3/4, and the config-2
kernel time at 4096 robots.  One JSON line per build; run once per build_ab/<name> on the GPU box."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, os.path.join("tests", "tools"))
from oracle import model as O
from parity_probe import make_ctx

name = sys.argv[1] if len(sys.argv) > 1 else "in-tree"
rec = {"build": name}
for cfg, seed in ((1, 0), (3, 1), (4, 2)):
    m = O.config_model(cfg, final_t=6.0)
    g = np.random.default_rng(seed).uniform(-1, 1, (8, m.genome_len()))
    steps = 300
    ctx = make_ctx(m, 32)
    ns = []
    for r in range(8):
        _, otv, _, _ = m.run(g[r], trace_steps=steps)
        _, tv, _, _ = ctx.trace(g[r], steps)
        err = np.abs(tv - otv).reshape(steps, -1).max(axis=1) / 3.0
        bad = np.nonzero(err > 1e-4)[0]
        ns.append(int(bad[0]) if len(bad) else steps)
    ctx.close()
    rec["config%d" % cfg] = {"N": ns, "N_min": min(ns), "N_median": float(np.median(ns))}
m = O.config_model(2, final_t=30.0)
g = np.random.default_rng(0).uniform(-1, 1, (4096, m.genome_len()))
ctx = make_ctx(m, 32)
ms = []
for _ in range(4):
    ctx.eval(g)
    ms.append(ctx.last_kernel_ms())
rec["config2_kernel_ms"] = float(np.min(ms[1:]))
rec["config2_voxel_steps_per_sec"] = 4096 * 10 * 1801 / (rec["config2_kernel_ms"] * 1e-3)
ctx.close()
print(json.dumps(rec))
