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


def build(cfg: int) -> dict:
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


if __name__ == "__main__":
    for cfg in SEEDS:
        np.savez_compressed(os.path.join(HERE, f"config{cfg}.npz"), **build(cfg))
        print("wrote config%d.npz" % cfg)
