"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/vsr_oracle.c, double precision).

The reference itself cannot produce vectors here (Java/Maven project, no JVM, 2dhmsr + dyn4j not on disk: SURVEY.md 8c), so these
are REGRESSION fixtures of the restated algorithm, not outputs of the reference: they freeze the oracle (tests/test_golden_cpu.py
fails if a later edit changes its numbers) and give the CUDA parity tests a committed target (tests/test_gpu_parity.py).
Per BASELINE config 1/3/4: 2 seeded genomes, the first 40 steps of voxel-polygon vertices and applied controls, and the fitness
and energy terms of a 3 s episode (transient 1 s).        python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import model as O  # noqa: E402

SEEDS = {1: 0, 3: 1, 4: 2}
STEPS, T = 40, 3.0


# the controller kinds added after the BASELINE configs: name -> (model, genome transform)
def extra_model(name: str):
    if name == "hetero":  # brainHeteroDist<funsGrid on config 4's robot and terrain
        return O.Model("comb-7x2", "steppy-0.25-1-0", "uniform-t+vxy+a-0", O.CTRL_MLP_DISTRIBUTED_HETERO, O.default_params(T), signals=1)
    if name == "timesin":  # TimeFunctions sinusoids on config 3's robot and terrain
        return O.Model("worm-8x2", "hilly-1-10-0", "", O.CTRL_TIME_SIN, O.default_params(T))
    raise ValueError(name)


def extra_genomes(name: str, m) -> np.ndarray:
    g = np.random.default_rng({"hetero": 7, "timesin": 8}[name]).uniform(-1, 1, (2, m.genome_len()))
    if name == "hetero":
        # small weights keep the MLP outputs away from +-1: a fully contracted voxel sits on the contraction-limit knife edge, where
        # 1e-13 rounding differences between two implementations of the same algorithm flip a rope within a few steps
        g = g * 0.15
    if name == "timesin":  # the decode of bodyBrainSin: frequency in [0.5, 2], phase in [0, pi], amplitude in [0, 1]
        g = g.reshape(2, -1, 3)
        g = np.stack([0.5 + 1.5 * (g[..., 0] + 1) / 2, np.pi * (g[..., 1] + 1) / 2, (g[..., 2] + 1) / 2], axis=-1).reshape(2, -1)
    return g


def build(cfg) -> dict:
    if isinstance(cfg, str):
        m = extra_model(cfg)
        g = extra_genomes(cfg, m)
    else:
        m = O.config_model(cfg, final_t=T)
        g = np.random.default_rng(SEEDS[cfg]).uniform(-1, 1, (2, m.genome_len()))
    tv, tc = [], []
    for r in range(2):
        _, v, _, c = m.run(g[r], trace_steps=STEPS)
        tv.append(v); tc.append(c)
    fit, res = m.eval_batch(g, n_threads=2, want_results=True)
    return dict(genomes=g, vertices=np.stack(tv), controls=np.stack(tc), fitness=fit,
                distance=np.array([r.distance for r in res]), time=np.array([r.time for r in res]),
                control_energy=np.array([r.control_energy for r in res]), area_ratio_abs=np.array([r.area_ratio_abs for r in res]),
                k_first=np.array([r.k_first for r in res]), k_last=np.array([r.k_last for r in res]), n_steps=np.array([r.n_steps for r in res]))


# what java/harness/DumpLocomotion.java (the real 2dhmsr Locomotion, run by a maintainer with the jars or by `bench.py --impl reference`
# when it finds a JVM) needs to replay the same robots: header `shape;sensorConfig;mapper;terrain;finalT`, one genome per line
REF_SPECS = {1: ("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat"),
             3: ("worm-8x2", "uniform-a-0", "brainPhaseVals;f=1", "hilly-1-10-0"),
             4: ("comb-7x2", "uniform-t+vxy+a-0", "brainHomoDist;s=1<mlp;r=0.65;nIL=1", "steppy-0.25-1-0")}


def write_reference_inputs():
    for cfg, (shape, sensors, mapper, terrain) in REF_SPECS.items():
        g = np.load(os.path.join(HERE, f"config{cfg}.npz"))["genomes"]
        with open(os.path.join(HERE, f"config{cfg}.genomes.csv"), "w") as f:
            f.write(f"{shape};{sensors};{mapper};{terrain};{T}\n")
            for row in g:
                f.write(",".join(repr(float(x)) for x in row) + "\n")
        print("wrote config%d.genomes.csv" % cfg)


if __name__ == "__main__":
    if sys.argv[1:] == ["--csv"]:
        write_reference_inputs()
        sys.exit(0)
    only = sys.argv[1:]
    for cfg in list(SEEDS) + ["hetero", "timesin"]:
        name = cfg if isinstance(cfg, str) else f"config{cfg}"
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **build(cfg))
        print("wrote %s.npz" % name)
