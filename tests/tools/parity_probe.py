"""Dev probe (GPU box): CUDA trace vs oracle trace for the BASELINE configs, both precisions."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import model as O
from vsrevolution_b200 import _ffi as F


def make_ctx(m, precision):
    e = F.env_default()
    e.final_t, e.transient_t, e.precision = m.params.final_t, m.params.transient_t, precision
    ctx = F.Context(e, m.tx, m.ty)
    sens = [F.SensorDesc(s.kind, s.axes, s.rotated, s.soft_norm, s.dom_min, s.dom_max, s.avg_window, s.noise_sigma) for s in m.sensors]
    ctx.body_set(m.mask, sens)
    layers = [m.ctrl.layers[i] for i in range(m.ctrl.n_layers)]
    ctx.controller_set(m.ctrl_kind, layers, m.signals if m.ctrl_kind == 2 else 0, m.freq, m.amp, m.input_activation)
    return ctx


def main():
    T = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    for cfg, seed in ((1, 0), (3, 1), (4, 2)):
        m = O.config_model(cfg, final_t=T)
        G = m.genome_len()
        g = np.random.default_rng(seed).uniform(-1, 1, (8, G))
        ns = O.episode_steps(T, 1.0)[0]
        ores, otv, otb, otc = m.run(g[0], trace_steps=ns)
        for prec in (64, 32):
            ctx = make_ctx(m, prec)
            assert ctx.genome_len == G
            t0 = time.time()
            res, tv, tb, tc = ctx.trace(g[0], ns)
            dt = time.time() - t0
            scale = np.abs(otv - otv[:, :, :, :].mean(axis=(1, 2), keepdims=True)).max() + 1e-9
            err = np.abs(tv - otv).max(axis=(1, 2, 3))
            rel = err / 3.0  # relative to the voxel side length
            first_bad = int(np.argmax(rel > 1e-4)) if (rel > 1e-4).any() else -1
            print(f"cfg{cfg} fp{prec}: steps={ns} trace {dt:.2f}s  max|dv| step1={err[0]:.2e} step60={err[min(59,ns-1)]:.2e} "
                  f"step300={err[min(299,ns-1)]:.2e} last={err[-1]:.2e}  first step rel>1e-4: {first_bad}  "
                  f"vel gpu={res['velocity']:.6f} oracle={ores.velocity:.6f} ovf={res['contact_overflow']}")
            fit, r = ctx.eval(g, want_results=True)
            ofit = m.eval_batch(g, n_threads=8)
            print(f"      batch fitness max|d|={np.abs(fit-ofit).max():.3e}  gpu={np.round(fit[:4],5)} oracle={np.round(ofit[:4],5)} kernel_ms={ctx.last_kernel_ms():.2f} info={ctx.kernel_info()}")
            ctx.close()


if __name__ == "__main__":
    main()
