"""Parity evidence (run on the GPU box; writes gpurun_out/parity_report.json, copied to profiles/r2_parity_report.json).

The checker is the CPU oracle (oracle/vsr_oracle.c) -- the builder's restatement of 2dhmsr 1.12.1 / dyn4j, PARITY UNPINNED against the
Java reference (no JVM, no jars, no golden vectors exist for it).  Sections:
  first_steps   leading steps N for which the CUDA voxel vertices stay within 1e-4 (relative to the voxel side) of the oracle's,
                FP64 and FP32 kernels, configs 1/3/4, 8 seeded robots each
  population    random-genome config 2 at full size (4096 robots x 30 s): Spearman, |d fitness| statistics
  locomoting    robots that actually move (where contacts and the knife-edge flips dominate):
                  config 2: the population after 20 generations of the device numGA loop (4096 robots x 30 s)
                  config 3: 4096 PhaseSin worms x 60 s on hilly terrain, all of them and those with |v| > 0.3
                  config 4: the full 16384 robots x 30 s on steppy terrain
                FP32 and FP64 kernels against the oracle, absolute and relative to |v|
usage: python tests/tools/parity_report.py [--quick]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import model as O
from parity_probe import make_ctx

QUICK = "--quick" in sys.argv
SIDE = 3.0
THREADS = os.cpu_count() or 8


def spearman(a, b):
    ra, rb = np.argsort(np.argsort(a)), np.argsort(np.argsort(b))
    return float(np.corrcoef(ra, rb)[0, 1])


def cmp(a, b):
    d = np.abs(a - b)
    rel = d / np.maximum(np.abs(b), 1e-9)
    return {"n": int(len(a)), "spearman": spearman(a, b) if len(a) > 2 else None, "median_abs_diff": float(np.median(d)),
            "mean_abs_diff": float(d.mean()), "p90_abs_diff": float(np.quantile(d, 0.9)), "max_abs_diff": float(d.max()),
            "median_rel_diff": float(np.median(rel)), "p90_rel_diff": float(np.quantile(rel, 0.9)),
            "frac_within_1e-3": float((d < 1e-3).mean()), "frac_within_1e-2": float((d < 1e-2).mean()),
            "frac_within_1e-1": float((d < 1e-1).mean()),
            "mean_ref": float(b.mean()), "std_ref": float(b.std()), "mean_abs_ref": float(np.abs(b).mean())}


def compare(m, g, label, movers=0.3):
    t0 = time.perf_counter()
    ref = m.eval_batch(g, n_threads=THREADS)
    osec = time.perf_counter() - t0
    rec = {"robots": int(len(g)), "oracle_seconds": osec, "oracle_threads": THREADS,
           "frac_moving": float((np.abs(ref) > movers).mean()), "movers_threshold": movers}
    fits = {}
    for prec in (32, 64):
        ctx = make_ctx(m, prec)
        fits[prec] = ctx.eval(g)
        rec["fp%d_overflows" % prec] = int(ctx.contact_overflows())
        ctx.close()
        rec["fp%d_vs_oracle" % prec] = cmp(fits[prec], ref)
        mv = np.abs(ref) > movers
        if mv.sum() > 2:
            rec["fp%d_vs_oracle_movers" % prec] = cmp(fits[prec][mv], ref[mv])
    rec["fp32_vs_fp64"] = cmp(fits[32], fits[64])
    print(label, json.dumps(rec)[:1500])
    sys.stdout.flush()
    return rec


def evolved_population(n_pop, gens, T, seed=0):
    """The population after `gens` generations of the device-resident numGA loop on config 2 (sorted best first)."""
    from vsrevolution_b200 import _ffi
    from vsrevolution_b200.locomotion import population_evaluator
    ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", T, 1.0)
    ga = _ffi.GeneticAlgorithm(ev.ctx, n_pop, seed=seed)
    ga.init()
    for _ in range(gens):
        ga.step()
    g, f = ga.population()
    ga.close()
    ev.ctx.close()
    return g, f


def main():
    out = {"checker": "oracle/vsr_oracle.c (double, builder's restatement; PARITY UNPINNED against 2dhmsr/dyn4j)",
           "tolerance": "max over voxel vertices of |cuda - oracle| <= 1e-4 * voxel side (3.0)", "first_steps": {}, "population": {},
           "locomoting": {}}
    for cfg, seed in ((1, 0), (3, 1), (4, 2)):
        m = O.config_model(cfg, final_t=6.0)
        g = np.random.default_rng(seed).uniform(-1, 1, (8, m.genome_len()))
        steps = 300
        rec = {}
        for prec in (64, 32):
            ctx = make_ctx(m, prec)
            ns = []
            for r in range(8):
                _, otv, _, _ = m.run(g[r], trace_steps=steps)
                _, tv, _, _ = ctx.trace(g[r], steps)
                err = np.abs(tv - otv).reshape(steps, -1).max(axis=1) / SIDE
                bad = np.nonzero(err > 1e-4)[0]
                ns.append(int(bad[0]) if len(bad) else steps)
            rec["fp%d" % prec] = {"N_per_robot": ns, "N_min": min(ns), "N_median": float(np.median(ns))}
            ctx.close()
        out["first_steps"]["config%d" % cfg] = rec
        print(cfg, rec)
        sys.stdout.flush()

    n2 = 512 if QUICK else 4096
    T2 = 10.0 if QUICK else 30.0
    m = O.config_model(2, final_t=T2)
    g = np.random.default_rng(0).uniform(-1, 1, (n2, m.genome_len()))
    out["population"] = compare(m, g, "config2 random")
    out["population"]["episode_s"] = T2

    ge, fe = evolved_population(n2, 20, T2)
    rec = compare(m, ge, "config2 evolved")
    rec.update({"generations": 20, "episode_s": T2, "device_best_fitness": float(fe[0]), "device_median_fitness": float(np.median(fe))})
    out["locomoting"]["config2_evolved"] = rec

    n3 = 512 if QUICK else 4096
    T3 = 20.0 if QUICK else 60.0
    m3 = O.config_model(3, final_t=T3)
    g3 = np.random.default_rng(1).uniform(-1, 1, (n3, m3.genome_len()))
    rec = compare(m3, g3, "config3 worms")
    rec["episode_s"] = T3
    out["locomoting"]["config3_worms"] = rec

    n4 = 512 if QUICK else 16384
    T4 = 10.0 if QUICK else 30.0
    m4 = O.config_model(4, final_t=T4)
    g4 = np.random.default_rng(2).uniform(-1, 1, (n4, m4.genome_len()))
    rec = compare(m4, g4, "config4 combs", movers=0.1)
    rec["episode_s"] = T4
    out["locomoting"]["config4_combs"] = rec

    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1)


if __name__ == "__main__":
    main()
