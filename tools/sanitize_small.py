"""Small invocations of every entry point / controller kind for compute-sanitizer (memcheck, racecheck) on the GPU box:
compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from vsrevolution_b200 import _ffi as F
from vsrevolution_b200.locomotion import population_evaluator

T = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
cases = [("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat"),
         ("biped-4x3", "uniform-a-0.01", "brainCentralized<mlp;r=0.65;nIL=1", "hilly-1-10-0"),
         ("worm-8x2", "", "brainPhaseVals;f=1", "hilly-1-10-0"),
         ("comb-7x2", "uniform-t+vxy+a-0", "brainHomoDist;s=1<mlp;r=0.65;nIL=1", "steppy-0.25-1-0"),
         ("comb-7x2", "uniform-t+vxy+a-0", "brainHeteroDist;s=1<funsGrid", "steppy-0.25-1-0"),
         ("tripod-5x3", "uniform-r+vxy+t-0", "brainCentralized<mlp;r=0.65;nIL=1", "uphill-10"),
         ("box-1x1", "", "brainPhaseVals;f=1", "flat")]
for shape, sens, mapper, terrain in cases:
    ev = population_evaluator(shape, sens, mapper, terrain, T, 0.25)
    g = np.random.default_rng(0).uniform(-1, 1, (37, ev.genome_len))
    fit = ev.evaluate(g)
    _, res = ev.ctx.eval(g, want_results=True)
    _, _, cen = ev.ctx.eval_centers(g[:5])
    ev.ctx.trace(g[0], 5)
    print(shape, mapper, terrain, "finite", bool(np.isfinite(fit).all()), "overflow", int((res["contact_overflow"] > 0).sum()), flush=True)
ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "hilly-1-10-rnd", T, 0.25)
print("rnd terrains", np.isfinite(ev.evaluate(np.random.default_rng(1).uniform(-1, 1, (9, ev.genome_len)))).all(), flush=True)
ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", T, 0.25)
for solver in (F.SOLVER_NUM_GA, F.SOLVER_ES):
    ga = F.GeneticAlgorithm(ev.ctx, 40, seed=1, solver=solver)
    ga.init(); ga.step(); ga.step()
    print("solver", solver, "best", ga.best()[1], flush=True)
    ga.close()
