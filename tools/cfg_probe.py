"""Kernel time and a fitness checksum of the BASELINE configs 2 / 3 / 4 for the library selected by VSR_LIB (A/B inside one gpurun call).
  python tools/cfg_probe.py <label> [workloads, e.g. 2,3,4] [robots for configs 3 / 4]"""
import hashlib, json, sys
import numpy as np
sys.path.insert(0, ".")
from bench import WORKLOADS
from vsrevolution_b200.locomotion import population_evaluator
label = sys.argv[1] if len(sys.argv) > 1 else ""
wls = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "2,3,4").split(",")]
big = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
rec = {"label": label}
for wl in wls:
    w = WORKLOADS[wl]
    n = 4096 if wl == 2 else big
    ev = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], w["T"], 1.0)
    g = np.random.default_rng(w["seed"]).uniform(-1, 1, (n, ev.genome_len))
    ms = []
    for _ in range(3 if wl == 2 else 2):
        fit = ev.ctx.eval(g)
        ms.append(ev.ctx.last_kernel_ms())
    rec[f"cfg{wl}_ms"] = round(float(min(ms[1:])), 2)
    rec[f"cfg{wl}_sha"] = hashlib.sha1(fit.tobytes()).hexdigest()[:10]
    rec[f"cfg{wl}_ovf"] = int(ev.ctx.contact_overflows())
    rec[f"cfg{wl}_vsps"] = float(f"{n * ev.n_voxels * 1801 * (2 if wl == 3 else 1) / (min(ms[1:]) * 1e-3):.4g}")
    ev.ctx.close()
print(json.dumps(rec))
