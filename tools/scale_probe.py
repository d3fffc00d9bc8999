"""Kernel time vs resident warps/SM (dev probe). usage: scale_probe.py shape robots..."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
shape = sys.argv[1]
os.environ["VSR_LAUNCH_SHAPE"] = shape
from vsrevolution_b200.locomotion import population_evaluator
T = float(os.environ.get("PROBE_T", "10"))
ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", T, 1.0)
info = ev.ctx.kernel_info()
for n in [int(a) for a in sys.argv[2:]]:
    g = np.random.default_rng(0).uniform(-1, 1, (n, ev.genome_len))
    ev.evaluate(g); ev.evaluate(g)
    ms = ev.ctx.last_kernel_ms()
    warps = -(-n // 3)
    print(f"shape {shape} regs {info['regs_per_thread']} thr/cta {info['threads_per_cta']} sync {os.environ.get('VSR_SYNC_MASK','1')}: robots {n:6d} warps {warps:5d} ({warps/148:5.2f}/SM) kernel {ms:8.2f} ms  -> {n*10*601/ms/1e3:.3e} vox-steps/s")
