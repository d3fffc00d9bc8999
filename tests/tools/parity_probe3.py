import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import model as O
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity_probe import make_ctx
T = 5.0
for cfg, seed, robots in ((1, 0, (5, 6, 7)), (3, 1, (0, 1, 2)), (4, 2, (0, 1))):
    m = O.config_model(cfg, final_t=T)
    g = np.random.default_rng(seed).uniform(-1, 1, (8, m.genome_len()))
    ns = O.episode_steps(T, 1.0)[0]
    ctx = make_ctx(m, 64)
    for r in robots:
        ores, otv, otb, otc = m.run(g[r], trace_steps=ns)
        res, tv, tb, tc = ctx.trace(g[r], ns)
        e = np.abs(tb - otb).max(axis=(1, 2))
        ks = [k for k in range(1, ns) if e[k] > 30 * e[k - 1] and e[k] > 1e-12]
        line = " ".join(f"{k+1}:{e[k-1]:.1e}->{e[k]:.1e}" for k in ks[:6])
        print(f"cfg{cfg} robot {r}: err@10={e[9]:.1e} @50={e[49]:.1e} @100={e[99]:.1e} @200={e[199]:.1e} @300={e[-1]:.1e} jumps[{line}] "
              f"ropes={ores.n_rope_vel} blocks={ores.n_block_vel} manifolds={ores.n_manifolds} detect={ores.n_detect_pairs}")
    ctx.close()
