#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: voxel-steps/s and 30 s-episode evals/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--robots R] [--workload 2|3|4]

One "step" = one pass of the hot path over one population batch: R robots (default 4096 per GPU, BASELINE
config 2: biped-4x3, CentralizedSensing 10-7-10 MLP, flat terrain, 30 s at dt=1/60 -> 1801 physics steps each).
N>1: one process per GPU (torchrun), robots sharded by index, no data-path collective ("weak" scaling: every
GPU gets its own R robots); timing is max over ranks.

`value`  : device-resident throughput (genomes already in HBM), CUDA events on the launch stream.
`e2e`    : same metric through PopulationEvaluator.evaluate() with HOST buffers (H2D + kernel + D2H).
`roofline`: FP32-issue roofline of the episode kernel: algorithmic flops from the ORACLE's
           operation tallies, live CUDA-event kernel time.
`cpu_baseline`: the CPU oracle (C port of the reference algorithm, "port") on the host cores, bounded sample.
`--impl reference`: times that CPU port alone on the SAME config (one full population batch per step when the host cores
           finish K of them within ~4.5 minutes, else the largest sample that does); it first probes for a JVM and the
           reference jars (`java_reference` key) -- the Java reference cannot run in the build image (no JVM, deps un-vendored).
`parity`  : a 64-robot sample of the workload evaluated by the CUDA path and by the oracle, outside the timed region; its
           `evolved_population` key repeats that on 64 robots of the population 20 `numGA` generations leave behind (robots that move),
           FP32 kernel and FP64 kernel against the oracle.
`other_configs`: one timed launch each of BASELINE configs 3 and 4 at full size (N=1 only).
`multi_device_exact` (N>1): rank 0 also drives ONE context over all N devices (in-process sharding with peer copies) and checks
           it bit for bit against a single-device context.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- algorithmic flop constants (counted on the ORACLE's source) -----------------
FLOPS = dict(spring_vel=38, spring_init=71, rope_vel=40, rope_init=30, weld_vel=78, weld_init=54, weld_pos=115,
             contact_pt_vel=47, contact_pt_pos=33, contact_init=43, detect_pair=280, body_step=48, voxel_act=140)

# DRAM bytes per launch from the committed ncu capture (profiles/r2_prof_a.csv: dram__bytes_read.sum + dram__bytes_write.sum;
# the 32 KB of fitness values stay in L2 within the capture): {(workload, robots): bytes}
NCU_DRAM_BYTES = {(2, 4096): 5306880 + 0}

WORKLOADS = {
    2: dict(name="config2: biped-4x3, brainCentralized<mlp;r=0.65;nIL=1 (10-7-10), uniform-a-0, flat, 30 s",
            shape="biped-4x3", sensors="uniform-a-0", mapper="brainCentralized<mlp;r=0.65;nIL=1", terrain="flat",
            T=30.0, seed=0, oracle_cfg=1),
    3: dict(name="config3: worm-8x2, brainPhaseVals;f=1, hilly-1-10-0, 60 s", shape="worm-8x2", sensors="",
            mapper="brainPhaseVals;f=1", terrain="hilly-1-10-0", T=60.0, seed=1, oracle_cfg=3),
    4: dict(name="config4: comb-7x2, brainHomoDist;s=1<mlp;r=0.65;nIL=1 (8-5-5), uniform-t+vxy+a-0, steppy-0.25-1-0, 30 s",
            shape="comb-7x2", sensors="uniform-t+vxy+a-0", mapper="brainHomoDist;s=1<mlp;r=0.65;nIL=1",
            terrain="steppy-0.25-1-0", T=30.0, seed=2, oracle_cfg=4),
}


def algorithmic_flops_per_robot(res, n_vox, n_springs, n_ropes, n_welds) -> float:
    """Flops of one episode from the oracle's operation tallies."""
    steps = res.n_steps
    f = FLOPS
    return (res.n_spring_vel * f["spring_vel"] + steps * n_springs * f["spring_init"]
            + res.n_rope_vel * f["rope_vel"] + steps * n_ropes * f["rope_init"] + res.n_rope_pos * f["rope_vel"]
            + res.n_weld_vel * f["weld_vel"] + steps * n_welds * f["weld_init"] + res.n_weld_pos * f["weld_pos"]
            + res.n_contact_pt_vel * f["contact_pt_vel"] + res.n_contact_pt_pos * f["contact_pt_pos"]
            + res.n_manifolds * 2 * f["contact_init"] + res.n_detect_pairs * f["detect_pair"]
            + steps * n_vox * 4 * f["body_step"] + steps * n_vox * f["voxel_act"])


def cpu_port(w, sample, threads):
    """Time the oracle on `sample` robots of workload w with `threads` host threads; returns (voxel-steps/s, flops/robot)."""
    from oracle import model as O
    m = O.config_model(w["oracle_cfg"], final_t=w["T"])
    g = np.random.default_rng(w["seed"]).uniform(-1, 1, (sample, m.genome_len()))
    O.lib()
    t0 = time.perf_counter()
    fit, results = m.eval_batch(g, n_threads=threads, want_results=True)
    dt = time.perf_counter() - t0
    import ctypes as C
    mask = m.mask
    cnt = [C.c_int() for _ in range(4)]
    O.lib().vo_counts(C.byref(m.params), m.w, m.h, mask.ctypes.data_as(C.c_void_p), *[C.byref(c) for c in cnt])
    nv, ns, nr, nw = [c.value for c in cnt]
    flops = float(np.mean([algorithmic_flops_per_robot(r, nv, ns, nr, nw) for r in results]))
    steps = results[0].n_steps
    return sample * nv * steps / dt, flops, steps, nv, dt


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peak_fp32_tflops():
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    mhz = peaks.get("sm_max_mhz", 1965.0)
    return 148 * 128 * 2 * mhz * 1e6 / 1e12, peaks.get("hbm_gbs", 6650.0), ("measured" if peaks else "fallback")


def probe_java_reference() -> dict:
    """SURVEY.md 8c-4 / BASELINE.md section 3: the only true parity pin is the real `new Locomotion(...).apply(robot)`
    (locomotion/Starter.java:87-93) of 2dhmsr 1.12.1.  Look for a JVM and the jars; never read /root/reference at run time."""
    import glob
    import shutil
    java = shutil.which("java")
    javac = shutil.which("javac")
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    jars = sorted(glob.glob(os.path.join(ref_dir, "**", "*.jar"), recursive=True))
    info = {"java": None, "javac": bool(javac), "jars": [os.path.relpath(j, ROOT) for j in jars][:8]}
    if java:
        try:
            r = subprocess.run([java, "-version"], capture_output=True, text=True, timeout=20)
            info["java"] = (r.stderr or r.stdout).strip().splitlines()[0] if (r.stderr or r.stdout).strip() else "present"
        except Exception as e:  # a broken launcher counts as absent
            info["java"] = None
            info["error"] = repr(e)
    need = [n for n, ok in (("JVM", info["java"]), ("javac", javac),
                            ("2dhmsr jar under baseline/_ref", any("TwoDimHighlyModularSoftRobots" in j or "2dhmsr" in j.lower() for j in jars))) if not ok]
    if need:
        info["status"] = "unavailable: missing " + ", ".join(need)
        return info
    # everything is there: run the committed harness (java/harness/DumpLocomotion.java) on the golden genomes; it writes
    # tests/golden/ref_*.npz-compatible CSVs that tests/tools/import_reference_traces.py turns into fixtures
    try:
        out = os.path.join(ROOT, "gpurun_out", "java_ref")
        os.makedirs(out, exist_ok=True)
        cp = os.pathsep.join(jars + [out])
        subprocess.check_call([javac, "-cp", cp, "-d", out, os.path.join(ROOT, "java", "harness", "DumpLocomotion.java")], timeout=300)
        subprocess.check_call([java, "-cp", cp, "DumpLocomotion", os.path.join(ROOT, "tests", "golden"), out], timeout=1800)
        info["status"] = "ran: traces under gpurun_out/java_ref (import with tests/tools/import_reference_traces.py)"
    except Exception as e:
        info["status"] = "unavailable: harness failed: " + repr(e)[:200]
    return info


def shared_config(w, robots, n_steps, nv, G):
    """The `config` object both arms print (so that the driver's same-config check compares like with like)."""
    return {"workload": w["name"], "robots_per_gpu": robots, "physics_steps": n_steps, "voxels": nv, "genome_len": G,
            "l2": "GPU arm: 256 MiB device write between timed iterations (inputs 5 MB << L2); CPU arm: not applicable"}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import model as O
    threads = os.cpu_count() or 1
    java = probe_java_reference()
    m = O.config_model(w["oracle_cfg"], final_t=w["T"])
    G = m.genome_len()
    # size the step: the full population batch of the GPU arm if K of them fit the time budget, else the largest sample that does
    probe_n = max(threads, min(args.robots, 8 * threads))
    _, _, steps, nv, probe_dt = cpu_port(w, probe_n, threads)
    per_robot = probe_dt / probe_n
    budget_s = float(os.environ.get("VSR_REF_BUDGET_S", "270"))
    sample = args.robots if args.ref_sample <= 0 else args.ref_sample
    fit_n = int(budget_s / max(1, args.steps) / per_robot)
    if sample > fit_n:
        sample = max(threads, (fit_n // threads) * threads)
    for _ in range(args.warmup):  # warm-up passes only page the library in and spin the pool up: small batches
        cpu_port(w, probe_n, threads)
    vals, tot = [], 0.0
    for _ in range(args.steps):
        v, flops, steps, nv, dt = cpu_port(w, sample, threads)
        vals.append(v); tot += dt
    value = float(np.mean(vals))
    full = sample == args.robots
    line = {"impl": "reference", "metric": "voxel-steps/sec", "value": value, "unit": "voxel-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(w, args.robots, steps, nv, G),
            "evals_per_sec": value / (nv * steps),
            "java_reference": java["status"], "java_probe": java,
            "cpu_baseline": {"value": value, "unit": "voxel-steps/s", "cores": threads, "kind": "port",
                             "sample": (f"the full {sample}-robot batch of the workload per step" if full else
                                        f"{sample} of the workload's {args.robots} robots per step (what {threads} cores finish in the time budget)")
                                       + ", C port of the reference algorithm (oracle/vsr_oracle.c, PARITY UNPINNED); the Java "
                                         "reference itself cannot run here (no JVM, 2dhmsr/dyn4j un-vendored): " + java["status"]},
            "e2e": {"value": value, "unit": "voxel-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL prints its version banner on stdout) must not break the one-JSON-line contract: everything written to
    fd 1 during the run goes to stderr; emit() writes the result line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def spearman(a, b):
    ra, rb = np.argsort(np.argsort(a)), np.argsort(np.argsort(b))
    return float(np.corrcoef(ra, rb)[0, 1])


def parity_block(ev, w, n=64):
    """Driver-visible parity evidence, OUTSIDE the timed region: the same `n` seeded robots of the workload through the CUDA
    path (vsr_eval) and through the oracle (checker only), full-length episodes; plus how many leading steps the FP32 kernel's
    voxel vertices stay within 1e-4 of the voxel side of the oracle's for the first 4 robots."""
    from oracle import model as O
    m = O.config_model(w["oracle_cfg"], final_t=w["T"])
    g = np.random.default_rng(w["seed"] + 77).uniform(-1, 1, (n, m.genome_len()))
    ref = m.eval_batch(g, n_threads=os.cpu_count() or 1)
    fit = ev.evaluate(g)
    d = np.abs(fit - ref)
    lead = []
    steps = 120
    for r in range(4):
        _, otv, _, _ = m.run(g[r], trace_steps=steps)
        _, tv, _, _ = ev.ctx.trace(g[r], steps)
        err = np.abs(tv - otv).reshape(steps, -1).max(axis=1) / 3.0
        bad = np.nonzero(err > 1e-4)[0]
        lead.append(int(bad[0]) if len(bad) else steps)
    return {"robots": n, "checker": "oracle/vsr_oracle.c (C restatement, double) -- PARITY UNPINNED against 2dhmsr itself",
            "spearman": spearman(fit, ref), "median_abs_dfitness": float(np.median(d)), "max_abs_dfitness": float(d.max()),
            "frac_within_1e-3": float((d < 1e-3).mean()), "fitness_std": float(ref.std()),
            "leading_steps_within_1e-4_of_voxel_side": lead, "tolerance": "north_star: 1e-4 relative over the first N steps; "
            "fitness statistically (the model has knife-edge discontinuities, DESIGN.md 6)"}


def parity_evolved_block(ev, w, n=64, generations=20, n_pop=4096):
    """Parity on robots that actually locomote: `n` robots spread over the population that `generations`
    generations of the device `numGA` loop leave behind, evaluated by the FP32 kernel, by the FP64 kernel (the chaotic floor: two
    roundings of one algorithm) and by the oracle (checker only).  Outside the timed region."""
    from oracle import model as O
    from vsrevolution_b200 import _ffi
    from vsrevolution_b200.locomotion import population_evaluator
    ga = _ffi.GeneticAlgorithm(ev.ctx, n_pop, seed=0)
    ga.init()
    for _ in range(generations):
        ga.step()
    pop, _ = ga.population()
    ga.close()
    g = np.ascontiguousarray(pop[np.linspace(0, n_pop - 1, n).astype(int)])
    m = O.config_model(w["oracle_cfg"], final_t=w["T"])
    ref = m.eval_batch(g, n_threads=os.cpu_count() or 1)
    fit = ev.evaluate(g)
    ev64 = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], w["T"], 1.0, precision=64)
    fit64 = ev64.evaluate(g)
    ev64.ctx.close()
    rel = lambda a: float(np.median(np.abs(a - ref) / np.maximum(np.abs(ref), 1e-9)))
    return {"robots": n, "generations": generations, "n_pop": n_pop, "mean_abs_velocity": float(np.abs(ref).mean()),
            "fp32_vs_oracle": {"spearman": spearman(fit, ref), "median_rel_dfitness": rel(fit)},
            "fp64_kernel_vs_oracle": {"spearman": spearman(fit64, ref), "median_rel_dfitness": rel(fit64)},
            "note": "robots moving at several voxel sides per second diverge chaotically within the 30 s episode in any two roundings of "
                    "the algorithm; the FP64 kernel against the oracle is that floor, the FP32 kernel sits on it (DESIGN.md 13.4)"}


def other_config(wk, robots, dev, torch):
    """One warm-up and ONE timed launch of another BASELINE config at full size (device-resident genomes, CUDA events)."""
    from vsrevolution_b200.locomotion import population_evaluator
    from vsrevolution_b200 import _ffi
    w = WORKLOADS[wk]
    ev = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], w["T"], 1.0, devices=[dev.index])
    n_steps = _ffi.episode_steps(w["T"], 1.0)[0]
    g = torch.from_numpy(np.random.default_rng(w["seed"]).uniform(-1, 1, (robots, ev.genome_len))).to(dev)
    fit = torch.empty(robots, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ev.ctx.eval_device(g.data_ptr(), min(robots, 2048), fit.data_ptr(), 0, st)  # warm-up: a sub-wave batch pages the kernel in
    torch.cuda.synchronize()
    ev.ctx.eval_device(g.data_ptr(), robots, fit.data_ptr(), 0, st)
    ms = ev.ctx.last_kernel_ms()
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(fit).all().item())
    out = {"workload": w["name"], "robots": robots, "physics_steps": n_steps, "voxels": ev.n_voxels, "kernel_ms": ms,
           "voxel_steps_per_sec": robots * ev.n_voxels * n_steps / (ms * 1e-3), "evals_per_sec": robots / (ms * 1e-3),
           "fitness_finite": ok, "contact_pool_overflows": int(ev.ctx.contact_overflows()), "kernel": ev.ctx.kernel_info()}
    ev.ctx.close()
    return out


def multi_device_check(w, world):
    """Rank 0 drives ONE context over all `world` devices (in-process sharding, peer copies in the generation loop) and compares
    it bit for bit with a single-device context: 101 robots (ragged over the devices), short episodes."""
    from vsrevolution_b200.locomotion import population_evaluator
    from vsrevolution_b200 import _ffi
    one = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], 3.0, 1.0, devices=[0])
    many = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], 3.0, 1.0, devices=list(range(world)))
    g = np.random.default_rng(5).uniform(-1, 1, (101, one.genome_len))
    a, b = one.evaluate(g), many.evaluate(g)
    exact = bool(np.array_equal(a, b)) and bool(np.isfinite(a).all())
    # the generation loop on the multi-device context (peer copies of genomes out and fitness back) against the single-device loop
    ga1 = _ffi.GeneticAlgorithm(one.ctx, 96, seed=4)
    gan = _ffi.GeneticAlgorithm(many.ctx, 96, seed=4)
    for ga in (ga1, gan):
        ga.init()
        ga.step()
    p1, f1 = ga1.population()
    pn, fn = gan.population()
    exact_ga = bool(np.array_equal(p1, pn)) and bool(np.array_equal(f1, fn))
    ga1.close(); gan.close()
    launches = int(many.kernel_launches())
    one.ctx.close(); many.ctx.close()
    return {"devices": world, "robots": 101, "eval_bit_identical": exact, "generation_loop_bit_identical": exact_ga,
            "episode_kernel_launches": launches}


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--robots", type=int, default=4096, help="robots per GPU per step")
    ap.add_argument("--workload", type=int, default=2, choices=sorted(WORKLOADS))
    ap.add_argument("--ref-sample", type=int, default=0, help="robots per reference step; 0 = the GPU arm's batch (--robots)")
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the 64-robot parity block")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the one-launch timings of configs 3 and 4 (N=1 only)")
    ap.add_argument("--ga-generations", type=int, default=3,
                    help="also time this many generations of the device-resident numGA loop (config 5); 0 = skip")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import torch
    import torch.distributed as dist
    from vsrevolution_b200.locomotion import population_evaluator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    ev = population_evaluator(w["shape"], w["sensors"], w["mapper"], w["terrain"], w["T"], 1.0, devices=[local])
    G, nv = ev.genome_len, ev.n_voxels
    n = args.robots
    # synthetic genomes U(-1,1) (UniformDoubleFactory(-1,1), builder/solver/DoublesStandard.java:52-53); rank-distinct
    host_g = np.random.default_rng(w["seed"] + 1000 * rank).uniform(-1, 1, (n, G))
    pinned_g = torch.from_numpy(host_g).pin_memory()
    host_fit = np.empty(n)
    d_g = torch.from_numpy(host_g).to(dev)
    d_fit = torch.empty(n, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    from vsrevolution_b200 import _ffi
    n_steps = _ffi.episode_steps(w["T"], 1.0)[0]

    def device_step():
        ev.ctx.eval_device(d_g.data_ptr(), n, d_fit.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    launches0 = ev.kernel_launches()
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- timed: device-resident -------------------------------------------------------------------
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    for a, b in evs:
        flush.fill_(1)  # L2 flush between timed iterations (256 MiB write)
        a.record()
        device_step()
        b.record()
        kernel_ms.append(ev.ctx.last_kernel_ms())  # CUDA events around the kernel on its launch stream; syncs
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- timed: end to end through the host API (pinned host genomes -> fitness on host) --------------
    pg = pinned_g.numpy()
    for _ in range(2):
        ev.ctx.eval_into(pg, host_fit)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev.ctx.eval_into(pg, host_fit)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ev.kernel_launches() - launches0
    # ---- config 5: the same evaluator inside the numGA generation loop (device-resident genomes) ---------
    ga_ms, ga_info = 0.0, None
    if args.ga_generations > 0:
        from vsrevolution_b200.solver import num_ga
        n_pop = n * world  # whole-job population; every rank evaluates its n robots and the ranks exchange 8 B per robot
        loop = num_ga().build({"nPop": str(n_pop), "nEval": "0"}).build(ev, seed=w["seed"], process_group=dist.group.WORLD if world > 1 else None)
        loop.init()
        loop.step()
        l0, g0 = ev.kernel_launches(), loop.ga.kernel_launches
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.ga_generations):
            loop.step()
        best = loop.ga.best_fitness()  # 8-byte D2H: the only thing that leaves the device
        ga_ms = (time.perf_counter() - t0) * 1e3
        ga_kernel_ms = ev.ctx.last_kernel_ms()  # episode kernel of the last generation (evolved genomes, not the random ones)
        barrier()
        ga_info = {"n_pop": n_pop, "generations": args.ga_generations, "best_fitness": best, "tournament": loop.ga.tournament_size,
                   "episode_kernel_ms": ga_kernel_ms,
                   "gpu_launches": int(ev.kernel_launches() - l0 + loop.ga.kernel_launches - g0)}
        launches += ga_info["gpu_launches"]
    t = torch.tensor([dev_ms, e2e_ms, ga_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, ga_ms = t.tolist()
    fit_ok = bool(np.isfinite(host_fit).all()) and bool(torch.isfinite(d_fit).all().item())
    overflows = int(ev.ctx.contact_overflows())  # this rank's warps whose contact pool overflowed in any launch above (0 = reliable)

    kinfo = ev.ctx.kernel_info()  # launch shape of the timed launches (before the extra blocks below launch anything else)
    # ---- outside the timed region: parity sample, the other configs (N=1), the in-process multi-device path (N>1) -------
    parity = others = multi = None

    def guarded(fn, *a):
        """The extra blocks must never cost the bench line itself: a failure is reported in place of the block."""
        try:
            return fn(*a)
        except Exception as e:  # noqa: BLE001
            return {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        if not args.no_parity:
            parity = guarded(parity_block, ev, w)
            if isinstance(parity, dict) and "error" not in parity and args.workload == 2:
                parity["evolved_population"] = guarded(parity_evolved_block, ev, w)
        if world > 1:
            multi = guarded(multi_device_check, w, world)
        elif not args.no_other_configs:
            others = [guarded(other_config, 3, 65536, dev, torch), guarded(other_config, 4, 16384, dev, torch)]
    if rank == 0:
        units = world * n * nv * n_steps * args.steps  # voxel-steps over all ranks and steps
        value = units / (dev_ms * 1e-3)
        e2e = units / (e2e_ms * 1e-3)
        line = {"metric": "voxel-steps/sec", "value": value, "unit": "voxel-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": shared_config(w, n, n_steps, nv, G),
                "parallelism": f"robots sharded by index over {world} GPU(s), one process per GPU, no collective on the data path",
                "kernel": kinfo,
                "evals_per_sec": value / (nv * n_steps),
                "e2e": {"value": e2e, "unit": "voxel-steps/s", "h2d_bytes_per_step": n * G * 8, "d2h_bytes_per_step": n * 8,
                        "evals_per_sec": e2e / (nv * n_steps), "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "fitness_finite": fit_ok, "contact_pool_overflows": overflows}
        if ga_info:
            per_gen = ga_ms / args.ga_generations
            ga_info.update({"ms_per_generation": per_gen, "evals_per_sec": ga_info["n_pop"] / (per_gen * 1e-3),
                            "voxel_steps_per_sec": ga_info["n_pop"] * nv * n_steps / (per_gen * 1e-3),
                            "loop_overhead_frac": max(0.0, 1.0 - ga_info["episode_kernel_ms"] / per_gen),
                            "what": "numGA = DoublesStandard(0.75,0.05,3,0.35): tournament + uniform crossover + Gaussian mutation + "
                                    "evaluation + (mu+lambda) truncation, genomes resident in HBM, wall clock incl. best-fitness D2H"})
            line["generation_loop"] = ga_info
        # roofline of the episode kernel (FP32 issue bound)
        peak, hbm_peak, src = peak_fp32_tflops()
        cpu = None
        flops_per_robot = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, flops_per_robot, _, _, cdt = cpu_port(w, args.cpu_sample, threads)
            cpu = {"value": v, "unit": "voxel-steps/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} robots of the same workload ({cdt:.1f} s wall), oracle/vsr_oracle.c double precision"}
        kms = float(np.mean(kernel_ms))
        if flops_per_robot:
            ach = flops_per_robot * n / (kms * 1e-3) / 1e12
            sust = (clocks or {}).get("sm_mhz") or None
            line["roofline"] = {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                "traffic": NCU_DRAM_BYTES.get((args.workload, n)), "kernel": "vsr_episode_kernel<float, ctrl, %d warps/CTA, %d CTA/SM>" % (kinfo["threads_per_cta"] // 32, kinfo["ctas_per_sm"]), "kernel_ms": kms,
                                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture profiles/r2_prof_a.csv; algorithmic bytes = genomes in + fitness out = %d" % (n * G * 8 + n * 8),
                                "units": {"flops": "algorithmic FP32 flop counted on the oracle (DESIGN.md 5)"},
                                "algorithmic_flops_per_voxel_step": flops_per_robot / (nv * n_steps),
                                "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz ({src} MEASURED_PEAKS.json)",
                                "frac_at_observed_clock": (ach / (148 * 128 * 2 * sust * 1e6 / 1e12)) if sust else None,
                                "hbm": {"achieved_gbs": (n * G * 8 + n * 8) / (kms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                        "note": "state is register/shared-memory resident; HBM carries genomes in, fitness out"}}
        line["cpu_baseline"] = cpu
        if parity:
            line["parity"] = parity
        if others:
            line["other_configs"] = others
        if multi:
            line["multi_device_exact"] = multi
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
