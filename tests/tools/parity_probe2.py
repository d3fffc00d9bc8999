import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import model as O
from vsrevolution_b200 import _ffi as F
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity_probe import make_ctx
np.set_printoptions(precision=3, linewidth=200)
T = 5.0
for cfg, seed in ((1, 0), (3, 1)):
    m = O.config_model(cfg, final_t=T)
    G = m.genome_len()
    g = np.random.default_rng(seed).uniform(-1, 1, (8, G))
    ns = O.episode_steps(T, 1.0)[0]
    for prec in (64, 32):
        ctx = make_ctx(m, prec)
        fit = ctx.eval(g)
        ofit = m.eval_batch(g, n_threads=8)
        print(f"cfg{cfg} fp{prec} batch diffs:", fit - ofit)
        # each robot alone via trace
        for r in (1, 2, 3):
            res, tv, tb, tc = ctx.trace(g[r], ns)
            print(f"   robot {r} alone: gpu {res['velocity']:.8f} batch {fit[r]:.8f} oracle {ofit[r]:.8f}")
        if prec == 32:
            ores, otv, otb, otc = m.run(g[0], trace_steps=ns)
            res, tv, tb, tc = ctx.trace(g[0], ns)
            ep = np.abs(tb[:, :, :2] - otb[:, :, :2]).max(axis=(1, 2))
            ea = np.abs(tb[:, :, 2] - otb[:, :, 2]).max(axis=1)
            ev = np.abs(tb[:, :, 3:5] - otb[:, :, 3:5]).max(axis=(1, 2))
            ew = np.abs(tb[:, :, 5] - otb[:, :, 5]).max(axis=1)
            ec = np.abs(tc - otc).max(axis=1)
            for k in list(range(0, 60, 2)) + list(range(60, ns, 20)):
                print(f"   step {k+1:4d} pos {ep[k]:.2e} ang {ea[k]:.2e} vel {ev[k]:.2e} w {ew[k]:.2e} ctrl {ec[k]:.2e}  ymin {otb[k,:,1].min():.3f}")
        ctx.close()
