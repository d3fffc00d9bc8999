"""Cost of the per-observation summaries (vsr_eval_ex, observation build of the kernel) next to the plain evaluation: config 2, 4096 robots x 30 s."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from vsrevolution_b200.locomotion import population_evaluator
ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 30.0, 1.0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = np.random.default_rng(0).uniform(-1, 1, (N, ev.genome_len))
out = {}
for name, fn in (("vsr_eval", lambda: ev.ctx.eval(g)),
                 ("vsr_eval_ex (no requests)", lambda: ev.ctx.eval_ex(g)),
                 ("vsr_eval_ex frozen masks", lambda: ev.ctx.eval_ex(g, frozen_masks=ev.ctx.broken_masks(0.3, 1, N))),
                 ("vsr_eval_ex posture 8x8", lambda: ev.ctx.eval_ex(g, posture_n=8)),
                 ("vsr_eval_ex posture + footprints 4,3 + centres", lambda: ev.ctx.eval_ex(g, posture_n=8, footprint_n=(4, 3), centers=True))):
    fn()
    t0 = time.perf_counter(); fn(); wall = time.perf_counter() - t0
    out[name] = {"kernel_ms": round(ev.ctx.last_kernel_ms(), 2), "wall_ms": round(wall * 1e3, 2), "info": ev.ctx.kernel_info()}
print(json.dumps(out, indent=1))
