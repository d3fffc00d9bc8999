"""compute-sanitizer targets of round 2 (small cases, every new code path): run as
   compute-sanitizer --tool memcheck python tools/sanitize_r2.py     (and --tool racecheck / synccheck / initcheck)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from vsrevolution_b200 import _ffi as F
from vsrevolution_b200.locomotion import population_evaluator

rng = np.random.default_rng(0)
# 1. the bench kernel (area-specialised, V2 position loop, libm math) on a few robots, short episode
ev = population_evaluator("biped-4x3", "uniform-a-0.01", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 1.5, 0.5)
g = rng.uniform(-1, 1, (7, ev.genome_len))
print("area kernel", ev.evaluate(g)[:3])
# 2. vsr_eval_ex: frozen masks + observation build (posture, footprints, centres)
fit, res, ex = ev.ctx.eval_ex(g, frozen_masks=ev.ctx.broken_masks(0.4, 3, 7), posture_n=8, footprint_n=(4, 3), centers=True)
print("eval_ex", fit[:3], int(ex["posture"].sum()), ex["centers"].shape)
# 3. per-evaluation terrains + staged inputs through the generation loop
rnd = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "hilly-1-10-rnd", 1.5, 0.5, transformation="broken-0.3-7")
from vsrevolution_b200.solver import num_ga
loop = num_ga().build({"nPop": "12", "nEval": "24"}).build(rnd, seed=1)
print("ga -rnd", loop.solve()[1])
# 4. per-voxel sensor lists with lidar / constants / cpg, centralized and distributed, steppy terrain (general narrowphase)
for shape, sensors, mapper, terrain in (("biped-4x3", "spinedTouchSighted-t-f-0.01", "brainCentralized<mlp;r=0.65;nIL=1", "steppy-0.25-1-0"),
                                        ("comb-7x2", "uniform-py+lf5+t-0", "brainHomoDist;s=1<mlp;r=0.65;nIL=1", "hilly-1-10-0"),
                                        ("worm-8x2", "", "brainPhaseVals;f=1", "downhill-30")):
    e2 = population_evaluator(shape, sensors, mapper, terrain, 1.5, 0.5)
    print(shape, sensors, e2.evaluate(0.2 * rng.uniform(-1, 1, (5, e2.genome_len)))[:2])
# 5. FP64 parity build + trace
e64 = population_evaluator("biped-4x3", "uniform-t+vxy+a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 1.0, 0.2, precision=64)
print("fp64", e64.evaluate(rng.uniform(-1, 1, (3, e64.genome_len)))[:2], e64.ctx.trace(rng.uniform(-1, 1, e64.genome_len), 5)[1].shape)
print("overflows", ev.ctx.contact_overflows())
