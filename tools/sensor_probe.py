"""Dev probe (GPU box): FP64 kernel vs oracle controls over the first steps, one sensor vocabulary at a time, to localise a mismatch."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import model as O
from vsrevolution_b200 import _ffi as F


def make_ctx(m, precision=64):
    e = F.env_default()
    e.final_t, e.transient_t, e.precision = m.params.final_t, m.params.transient_t, precision
    ctx = F.Context(e, m.tx, m.ty)
    conv = lambda s: F.SensorDesc(s.kind, s.axes, s.rotated, s.soft_norm, s.dom_min, s.dom_max, s.avg_window, s.noise_sigma, s.param, s.n_rays, s.side)
    ctx.body_set_voxels(m.mask, [conv(s) for s in m.vox_sensors], m.vox_off)
    layers = [m.ctrl.layers[i] for i in range(m.ctrl.n_layers)]
    ctx.controller_set(m.ctrl_kind, layers, 0, m.freq, m.amp, m.input_activation)
    return ctx


for terrain in ("flat", "hilly-1-10-0"):
    for sensors in sys.argv[1:] or ("uniform-lf5-0", "uniform-l5-0", "uniform-l1-0", "uniform-cpg-0", "uniform-px+py-0", "uniform-a-0",
                                    "spinedTouch-f-f-0", "spinedTouch-t-f-0", "spinedTouchSighted-f-f-0"):
        m = O.Model("biped-4x3", terrain, sensors, O.CTRL_MLP_CENTRALIZED, O.default_params(1.0, 0.2))
        g = np.random.default_rng(12).uniform(-1, 1, m.genome_len())
        _, otv, otb, otc = m.run(g, trace_steps=20)
        ctx = make_ctx(m)
        _, tv, tb, tc = ctx.trace(g, 20)
        ctx.close()
        ec = np.abs(tc - otc).max(axis=1)
        print(f"{terrain:14s} {sensors:26s} ctrl err step1 {ec[0]:.2e} step5 {ec[4]:.2e} step20 {ec[19]:.2e}  body err20 {np.abs(tb - otb).max():.2e}")
