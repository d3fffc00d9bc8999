"""Turn the CSVs written by java/harness/DumpLocomotion.java (the REAL 2dhmsr 1.12.1 Locomotion) into golden fixtures and report how far
the oracle -- and with it every [U] constant of DESIGN.md section 2 -- is from the reference.  This is the procedure that PINS parity
(SURVEY.md 8c-4); until somebody with a JVM and the jars runs it, parity stays "unpinned".

  java -cp <2dhmsr, dyn4j, JGEA, VSREvolution jars>:java/harness DumpLocomotion tests/golden gpurun_out/java_ref
  python tests/tools/import_reference_traces.py gpurun_out/java_ref        -> tests/golden/ref_config{1,3,4}.npz + a report

tests/test_golden_cpu.py::test_oracle_against_reference_traces picks the ref_*.npz files up when they exist."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as O  # noqa: E402


def load(path):
    rows, summary = [], {}
    for line in open(path):
        if line.startswith("# robot"):
            p = line.split()
            summary[int(p[2])] = dict(velocity=float(p[4]), distance=float(p[6]), time=float(p[8]))
        elif line[0].isdigit():
            rows.append([float(x) for x in line.split(",")])
    a = np.array(rows)
    n_robots = int(a[:, 0].max()) + 1
    n_vox, n_vert = int(a[:, 2].max()) + 1, int(a[:, 3].max()) + 1
    ts = np.unique(a[a[:, 0] == 0][:, 1])
    v = np.full((n_robots, len(ts), n_vox, n_vert, 2), np.nan)
    t_index = {t: i for i, t in enumerate(ts)}
    for r, t, vx, k, x, y in a:
        v[int(r), t_index[t], int(vx), int(k)] = (x, y)
    return ts, v, summary


def main(src):
    for cfg in (1, 3, 4):
        f = os.path.join(src, f"config{cfg}.ref.csv")
        if not os.path.exists(f):
            print("missing", f)
            continue
        ts, ref_v, summary = load(f)
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"config{cfg}.npz"))
        m = O.config_model(cfg, final_t=3.0)
        steps = min(len(ts), 120)
        errs = []
        for r in range(ref_v.shape[0]):
            _, otv, _, _ = m.run(gold["genomes"][r], trace_steps=steps)
            # the reference's vertex order inside a polygon may differ from the oracle's (NW NE SE SW): compare as point sets
            e = np.array([min(np.abs(np.roll(ref_v[r, k], s, axis=1) - otv[k]).max() for s in range(ref_v.shape[3])) for k in range(steps)])
            errs.append(e)
            bad = np.nonzero(e / 3.0 > 1e-4)[0]
            print(f"config{cfg} robot {r}: oracle within 1e-4 of the voxel side of the REFERENCE for the first {int(bad[0]) if len(bad) else steps} steps; "
                  f"step-1 error {e[0]:.3e}; reference velocity {summary[r]['velocity']:.6f} vs oracle {gold['fitness'][r]:.6f}")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"ref_config{cfg}.npz"), t=ts, vertices=ref_v,
                            velocity=np.array([summary[r]["velocity"] for r in sorted(summary)]), genomes=gold["genomes"][:ref_v.shape[0]])
        print("wrote tests/golden/ref_config%d.npz" % cfg)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "java_ref"))
