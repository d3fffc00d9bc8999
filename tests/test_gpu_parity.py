"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (stated here, as north_star asks):
  * FP64 kernel build vs oracle : body state within 1e-9 (absolute, world units) for the first 40 steps -- same
    algorithm, same constraint order; differences are rounding only.
  * FP32 product kernel vs oracle: voxel vertices within 1e-4 RELATIVE to the voxel side (3.0) for the first
    N >= 15 steps; N (the first step beyond the bound) is reported, it is where the first discrete solver
    decision (position-solver early exit / contraction-limit activation) flips between FP32 and FP64.
  * population fitness: Spearman rank correlation and median |d| over a seeded population, with the FP64-kernel
    vs oracle figure next to it (the chaotic floor of the model).
  * integer layout: bit-exact.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vsrevolution_b200 import _ffi
    _ffi.lib()
    return _ffi


def make_ctx(F, m, precision=32, devices=None, gravity=None):
    e = F.env_default()
    e.final_t, e.transient_t, e.precision = m.params.final_t, m.params.transient_t, precision
    if gravity is not None:
        e.gravity_y = gravity
    if devices:
        e.n_devices = len(devices)
        for i, d in enumerate(devices):
            e.devices[i] = d
    ctx = F.Context(e, m.tx, m.ty)
    conv = lambda s: F.SensorDesc(s.kind, s.axes, s.rotated, s.soft_norm, s.dom_min, s.dom_max, s.avg_window, s.noise_sigma, s.param, s.n_rays, s.side)
    if m.uniform:
        ctx.body_set(m.mask, [conv(s) for s in m.sensors])
    else:  # every voxel its own sensor list (spinedTouch, px / py / l5 ...)
        ctx.body_set_voxels(m.mask, [conv(s) for s in m.vox_sensors], m.vox_off)
    layers = [m.ctrl.layers[i] for i in range(m.ctrl.n_layers)]
    ctx.controller_set(m.ctrl_kind, layers, m.signals if m.ctrl_kind in (2, 3) else 0, m.freq, m.amp, m.input_activation)
    return ctx


def spearman(a, b):
    ra, rb = np.argsort(np.argsort(a)), np.argsort(np.argsort(b))
    return float(np.corrcoef(ra, rb)[0, 1])


CFG_SEED = {1: 0, 3: 1, 4: 2}


@pytest.mark.parametrize("cfg", [1, 3, 4])
def test_fp64_kernel_is_the_oracle_algorithm(F, oracle, cfg):
    m = oracle.config_model(cfg, final_t=3.0)
    g = np.random.default_rng(CFG_SEED[cfg]).uniform(-1, 1, (4, m.genome_len()))
    ctx = make_ctx(F, m, 64)
    for r in range(4):
        _, otv, otb, otc = m.run(g[r], trace_steps=40)
        res, tv, tb, tc = ctx.trace(g[r], 40)
        assert np.abs(tb - otb).max() < 1e-9, (cfg, r, np.abs(tb - otb).max(axis=(1, 2)))
        assert np.abs(tv - otv).max() < 1e-9 and np.abs(tc - otc).max() < 1e-9
        assert res["contact_overflow"] == 0
    ctx.close()


@pytest.mark.parametrize("cfg", [1, 3, 4])
def test_fp32_kernel_tracks_oracle_until_first_discrete_flip(F, oracle, bounds, cfg):
    m = oracle.config_model(cfg, final_t=3.0)
    g = np.random.default_rng(CFG_SEED[cfg]).uniform(-1, 1, (4, m.genome_len()))
    ctx = make_ctx(F, m, 32)
    firsts = []
    for r in range(4):
        _, otv, otb, otc = m.run(g[r], trace_steps=120)
        res, tv, tb, tc = ctx.trace(g[r], 120)
        rel = np.abs(tv - otv).max(axis=(1, 2, 3)) / 3.0
        bad = np.nonzero(rel > 1e-4)[0]
        n_ok = int(bad[0]) if len(bad) else 120
        firsts.append(n_ok)
        bounds.check(f"track32.cfg{cfg}.rel_err_first10", rel[:10].max(), 2e-6)      # while tracking, the error is FP32 rounding, ~1e-7
        bounds.check(f"track32.cfg{cfg}.ctrl_err_first10", np.abs(tc[:10] - otc[:10]).max(), 1e-5)  # control signals too
    print(f"cfg{cfg}: steps tracked within 1e-4 relative (FP32 vs FP64 oracle): {firsts}")
    bounds.check(f"track32.cfg{cfg}.N_min", min(firsts), 10, kind="min_int")        # floor = the measured minimum of these seeded robots
    ctx.close()


def test_population_fitness_agrees_statistically(F, oracle, bounds):
    """Config-2 population (shorter episode): FP32 kernel vs oracle, with FP64 kernel vs oracle as the floor."""
    m = oracle.config_model(2, final_t=10.0)
    n = 384
    g = np.random.default_rng(0).uniform(-1, 1, (n, m.genome_len()))
    ref = m.eval_batch(g, n_threads=8)
    out = {}
    for prec in (64, 32):
        ctx = make_ctx(F, m, prec)
        fit, res = ctx.eval(g, want_results=True)
        assert np.isfinite(fit).all() and (res["contact_overflow"] == 0).all()
        assert (res["n_steps"] == 600).all() and (res["k_first"] == 60).all() and (res["k_last"] == 599).all()
        out[prec] = (spearman(fit, ref), float(np.median(np.abs(fit - ref))), float(np.mean(np.abs(fit - ref) < 1e-3)))
        ctx.close()
    print("population parity (spearman, median|d|, frac |d|<1e-3): fp64", out[64], "fp32", out[32])
    bounds.check("pop2.fp64.spearman", out[64][0], 0.95, kind="spearman")
    bounds.check("pop2.fp32.spearman", out[32][0], 0.90, kind="spearman")
    bounds.check("pop2.fp32.median_abs", out[32][1], 1e-3)
    bounds.check("pop2.fp64.median_abs", out[64][1] + 1e-15, 1e-3)
    # distribution-level agreement of the fitness values
    bounds.check("pop2.fp32.mean_diff", abs(np.mean(fit) - np.mean(ref)), 2e-3)
    bounds.check("pop2.fp32.std_diff", abs(np.std(fit) - np.std(ref)), 3e-3)


def test_device_integer_layout_bit_exact(F, oracle):
    m = oracle.config_model(4, final_t=2.0)
    ctx = make_ctx(F, m)
    assert ctx.genome_len == 75 and ctx.n_voxels == 11
    assert ctx.counts() == (11, 198, 44, 20)
    cells = ctx.voxel_cells()
    assert cells.tolist() == [i for i, v in enumerate(m.mask.reshape(-1)) if v]
    cell, sensor, reading = ctx.input_map()
    assert len(cell) == 11 * 4 and cell[:4].tolist() == [0] * 4 and sensor[:4].tolist() == [0, 1, 1, 2] and reading[:4].tolist() == [0, 0, 1, 0]
    ctx.close()
    m1 = oracle.config_model(1, final_t=2.0)
    c1 = make_ctx(F, m1)
    assert c1.genome_len == 157 and c1.counts() == (10, 180, 40, 24)
    c1.close()


def test_edge_cases_and_errors(F, oracle):
    m = oracle.config_model(1, final_t=2.0)
    ctx = make_ctx(F, m)
    G = ctx.genome_len
    assert ctx.eval(np.zeros((0, G))).shape == (0,)                      # empty population
    rng = np.random.default_rng(5)
    g = rng.uniform(-1, 1, (13, G))
    full = ctx.eval(g)
    for n in (1, 2, 3, 4, 5, 12, 13):                                      # ragged: not a multiple of robots per warp / CTA
        assert np.array_equal(ctx.eval(g[:n]), full[:n]), n
    with pytest.raises(ValueError):
        ctx.eval(np.zeros((3, G - 1)))
    with pytest.raises(F.VsrIllegalArgument):                              # MLP I/O does not match the body
        ctx.controller_set(F.CTRL_MLP_CENTRALIZED, [9, 7, 10])
    ctx.controller_set(F.CTRL_MLP_CENTRALIZED, [10, 7, 10])
    big = np.ones((6, 6), dtype=np.uint8)
    with pytest.raises(F.VsrError) as ei:
        ctx.body_set(big, [])
    assert ei.value.code == F.VSR_EUNSUPPORTED
    ctx.close()
    # a window [transient_t, final_t) holding ONE observation has distance 0 over time 0: rejected, not NaN for every robot
    m1 = oracle.config_model(1, final_t=2.0)
    m1.params.transient_t = 1.995
    n_steps, k_first, k_last = F.episode_steps(2.0, m1.params.transient_t)
    assert k_first == k_last
    with pytest.raises(F.VsrIllegalArgument):
        make_ctx(F, m1)


def test_zero_gravity_zero_amplitude_robot_does_not_move(F, oracle):
    p = oracle.default_params(2.0, 0.5)
    m = oracle.Model("worm-8x2", "flat", "", oracle.CTRL_PHASE_SIN, p, amp=0.0)
    ctx = make_ctx(F, m, 32, gravity=0.0)
    res, tv, tb, tc = ctx.trace(np.zeros(16), 60)
    # FP32 + MUFU rsqrt: rest lengths are reproduced to ~1e-7, so residual velocities stay at rounding level
    assert np.abs(tb[:, :, 3:]).max() < 1e-5 and np.abs(tb[-1, :, :3] - tb[0, :, :3]).max() < 1e-5
    assert abs(res["velocity"]) < 1e-6
    ctx.close()


def test_full_size_batch_properties(F, oracle, bounds):
    """BASELINE config 2 at full size (4096 robots, 30 s): size-independent properties."""
    m = oracle.config_model(2)
    ctx = make_ctx(F, m)
    n, G = 4096, ctx.genome_len
    g = np.random.default_rng(0).uniform(-1, 1, (n, G))
    a, res = ctx.eval(g, want_results=True)
    assert np.isfinite(a).all() and (res["contact_overflow"] == 0).all()
    assert (res["n_steps"] == 1801).all() and (res["k_first"] == 60).all() and (res["k_last"] == 1800).all()
    np.testing.assert_allclose(res["time"], 1740 / 60.0, rtol=1e-12)
    np.testing.assert_allclose(res["velocity"] * res["time"], res["distance"], rtol=1e-12, atol=1e-15)
    assert np.array_equal(ctx.eval(g), a)                                   # deterministic
    perm = np.random.default_rng(1).permutation(n)
    assert np.array_equal(ctx.eval(g[perm]), a[perm])                       # robots are independent: permutation equivariant
    for i in (0, 1, 2, 4095):                                               # a robot alone == the same robot in the batch
        assert ctx.eval(g[i:i + 1])[0] == a[i]
    # sample of the population against the oracle
    idx = np.arange(0, n, 64)
    ref = m.eval_batch(g[idx], n_threads=8)
    bounds.check("full2.sample64.spearman", spearman(a[idx], ref), 0.8, kind="spearman")
    bounds.check("full2.sample64.median_abs", np.median(np.abs(a[idx] - ref)), 2e-3)
    assert abs(np.abs(a).max()) < 5.0                                       # biped top speed is a few units/s at most
    ctx.close()


def test_noisy_sensors_replay_java_random(F, oracle):
    """The reference's default sensor config `uniform-a-0.01` (Starter.java:152): Noisy(sensor, 0.01, seed 0).
    Oracle (C java.util.Random) and library (C++ java.util.Random, table on the device) are independent replays."""
    p = oracle.default_params(3.0, 1.0)
    m = oracle.Model("biped-4x3", "flat", "uniform-a-0.01", oracle.CTRL_MLP_CENTRALIZED, p)
    clean = oracle.Model("biped-4x3", "flat", "uniform-a-0", oracle.CTRL_MLP_CENTRALIZED, p)
    g = np.random.default_rng(0).uniform(-1, 1, (3, m.genome_len()))
    ctx = make_ctx(F, m, 64)
    for r in range(3):
        _, otv, otb, otc = m.run(g[r], trace_steps=40)
        _, _, _, ctc = clean.run(g[r], trace_steps=40)
        res, tv, tb, tc = ctx.trace(g[r], 40)
        assert np.abs(tb - otb).max() < 1e-9 and np.abs(tc - otc).max() < 1e-9
        assert np.abs(otc - ctc).max() > 1e-4  # the noise really reaches the controller
    ctx.close()
    ctx32 = make_ctx(F, m, 32)
    _, otv, _, otc = m.run(g[0], trace_steps=10)
    _, tv, _, tc = ctx32.trace(g[0], 10)
    assert np.abs(tv - otv).max() / 3.0 < 1e-4 and np.abs(tc - otc).max() < 1e-5
    ctx32.close()


def test_steppy_terrain_never_overflows_the_contact_pool(F, oracle, bounds):
    """Config 4 terrain (narrow steps): a corner mass can touch 3 polygons; the per-warp pool must hold them all."""
    m = oracle.config_model(4, final_t=10.0)
    ctx = make_ctx(F, m)
    g = np.random.default_rng(2).uniform(-1, 1, (4096, ctx.genome_len))
    fit, res = ctx.eval(g, want_results=True)
    assert np.isfinite(fit).all()
    assert int((res["contact_overflow"] > 0).sum()) == 0
    idx = np.arange(0, 4096, 128)
    ref = m.eval_batch(g[idx], n_threads=8)
    bounds.check("steppy4.sample32.median_abs", np.median(np.abs(fit[idx] - ref)), 2e-3)
    bounds.check("steppy4.sample32.spearman", spearman(fit[idx], ref), 0.5, kind="spearman")
    ctx.close()


def test_eval_device_equals_host_path(F, oracle):
    import torch
    m = oracle.config_model(3, final_t=3.0)
    ctx = make_ctx(F, m)
    g = np.random.default_rng(3).uniform(-1, 1, (70, ctx.genome_len))
    host = ctx.eval(g)
    dg = torch.from_numpy(g).cuda()
    dfit = torch.zeros(70, dtype=torch.float64, device="cuda")
    k0 = ctx.kernel_launches()
    ctx.eval_device(dg.data_ptr(), 70, dfit.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert ctx.kernel_launches() == k0 + 1 and ctx.last_kernel_ms() > 0
    assert np.array_equal(dfit.cpu().numpy(), host)
    ctx.close()


def test_host_mirror_end_to_end(F, oracle, bounds):
    """The reference-shaped call: mapper(genome) -> Robot -> Locomotion.apply_all -> Outcome.getVelocity."""
    from vsrevolution_b200.builder import build_mapper
    from vsrevolution_b200.locomotion import build_locomotion_task, population_evaluator
    from vsrevolution_b200.robot import Robot, RobotUtils
    T = 3.0
    target = Robot(None, RobotUtils.build_sensorizing_function("uniform-a-0")(RobotUtils.build_shape("biped-4x3")))
    mapper = build_mapper("brainCentralized<mlp;r=0.65;nIL=1").build_for(target)
    g = np.random.default_rng(0).uniform(-1, 1, (6, 157))
    task = build_locomotion_task("flat", T, 1.0)
    outs = task.apply_all([mapper(row) for row in g])
    ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", T, 1.0)
    fit = ev.evaluate(g)
    assert [o.get_velocity() for o in outs] == fit.tolist(), (outs, fit)
    assert task.apply(mapper(g[2])).sub_outcome(1.0, T).get_velocity() == fit[2]
    m = oracle.config_model(1, final_t=T)
    bounds.check("mirror.max_abs_vs_oracle", np.abs(fit - m.eval_batch(g)).max(), 3e-2)  # 2 s window after the first FP32-vs-FP64 solver flips
    # Starter.validate (Starter.java:122-135): the same robots on the validation terrains (default `flat,downhill-30`, :171)
    from vsrevolution_b200.locomotion import validate
    robots = [mapper(row) for row in g]
    vflat = validate(robots, "flat", "identity", 0, T, 1.0)
    assert [v.outcome.get_velocity() for v in vflat] == fit.tolist() and vflat[0].terrain_name == "flat" and vflat[0].seed == 0
    vdown = validate(robots, "downhill-30", "identity", 0, T, 1.0)
    assert np.mean([v.outcome.get_velocity() for v in vdown]) > np.mean(fit) + 0.5  # everybody slides down a 30 degree slope
    vrnd = validate(robots, "hilly-1-10-rnd", "identity", 3, T, 1.0)
    assert len({round(v.outcome.get_velocity(), 9) for v in vrnd}) == len(robots)
    with pytest.raises(F.VsrError):
        validate(robots, "flat", "broken-0.5", 0, T, 1.0)


def test_multi_device_sharding_is_exact(F, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    m = oracle.config_model(1, final_t=2.0)
    g = np.random.default_rng(9).uniform(-1, 1, (101, m.genome_len()))
    c1 = make_ctx(F, m, devices=[0])
    c2 = make_ctx(F, m, devices=[0, 1])
    assert np.array_equal(c1.eval(g), c2.eval(g))
    c1.close(); c2.close()


@pytest.mark.parametrize("cfg", [1, 3, 4])
def test_outcome_energy_terms_match_oracle(F, oracle, bounds, cfg):
    """vsr_result energy sums (inputs of Outcome.getControlPower / getAreaRatioPower / getCorrectedEfficiency,
    NamedFunctions.java:135-137) against the oracle on a window short enough to precede the first discrete divergence:
    FP64 kernel 1e-8 relative (same algorithm, same order); FP32 kernel 5e-2 relative (sum |da| accumulates the FP32
    trajectory differences of every voxel and step; the FP32 build tracks the oracle to 1e-4 only for the first N steps)."""
    m = oracle.config_model(cfg, final_t=0.7)
    m.params.transient_t = 0.2
    g = np.random.default_rng(CFG_SEED[cfg]).uniform(-1, 1, (6, m.genome_len()))
    _, ores = m.eval_batch(g, n_threads=4, want_results=True)
    for prec, tol in ((64, 1e-8), (32, 5e-2)):
        ctx = make_ctx(F, m, prec)
        _, res = ctx.eval(g, want_results=True)
        for i in range(len(g)):
            assert res["k_first"][i] == ores[i].k_first and res["k_last"][i] == ores[i].k_last
            for name in ("control_energy", "area_ratio_abs", "area_ratio_sq"):
                a, b = float(res[name][i]), getattr(ores[i], name)
                bounds.check(f"energy.cfg{cfg}.fp{prec}.{name}.rel", abs(a - b) / max(abs(b), 1e-3), tol)
            assert res["control_energy"][i] > 0 and res["area_ratio_abs"][i] > 0
        # the fitness-only call must give the same fitness (the energy path is results-mode only)
        assert np.array_equal(ctx.eval(g), res["velocity"])
        ctx.close()
    # host mirror getters
    from vsrevolution_b200.locomotion import Outcome
    o = Outcome(0.5, 5.0, 10.0, 0.0, 5.0, 1.0, 11.0, control_energy=4.0, area_ratio_energy=2.0)
    assert o.get_control_power() == 0.4 and o.get_area_ratio_power() == 0.2 and o.get_corrected_efficiency() == 1.0


@pytest.mark.parametrize("cfg", [1, 3, 4])
def test_cuda_path_against_committed_golden_vectors(F, oracle, cfg):
    """tests/golden/config*.npz (frozen oracle output): FP64 kernel vertices within 1e-9 and controls within 1e-9 for the 40
    recorded steps; FP32 kernel within 1e-4 of the voxel side for the first 10 steps."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"config{cfg}.npz"))
    m = oracle.config_model(cfg, final_t=3.0)
    steps = gold["vertices"].shape[1]
    for prec in (64, 32):
        ctx = make_ctx(F, m, prec)
        for r in range(2):
            _, tv, _, tc = ctx.trace(gold["genomes"][r], steps)
            if prec == 64:
                assert np.abs(tv - gold["vertices"][r]).max() < 1e-9 and np.abs(tc - gold["controls"][r]).max() < 1e-9
            else:
                assert np.abs(tv[:10] - gold["vertices"][r][:10]).max() < 1e-4 * 3.0
        if prec == 64:
            fit, res = ctx.eval(gold["genomes"], want_results=True)
            assert (res["k_first"] == gold["k_first"]).all() and (res["k_last"] == gold["k_last"]).all()
            assert np.abs(res["time"] - gold["time"]).max() < 1e-12
        ctx.close()


@pytest.mark.parametrize("name", ["hetero", "timesin"])
def test_cuda_path_against_golden_vectors_of_the_later_controller_kinds(F, oracle, name):
    """tests/golden/{hetero,timesin}.npz: FP64 kernel 1e-9 on the 40 recorded steps, FP32 kernel 1e-4 of the voxel side on the first 10."""
    import importlib.util, os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    gold = np.load(os.path.join(here, "golden", f"{name}.npz"))
    m = mg.extra_model(name)
    steps = gold["vertices"].shape[1]
    for prec in (64, 32):
        ctx = make_ctx(F, m, prec)
        for r in range(2):
            _, tv, _, tc = ctx.trace(gold["genomes"][r], steps)
            if prec == 64:
                err = np.abs(tv - gold["vertices"][r]).max(axis=(1, 2, 3))
                assert err.max() < 1e-9 and np.abs(tc - gold["controls"][r]).max() < 1e-9, (name, r, err)
            else:
                assert np.abs(tv[:10] - gold["vertices"][r][:10]).max() < 1e-4 * 3.0
        ctx.close()


def test_per_evaluation_terrains_match_single_terrain_runs(F, oracle, bounds):
    """vsr_eval_terrains (`-rnd` terrains, Starter.java:89-93): robot i on terrain i in ONE launch == robot i evaluated alone in a
    context built on terrain i (bit-identical, FP32), and the FP64 kernel tracks the oracle run on that terrain."""
    names = ["hilly-1-10-%d" % s for s in (3, 1234, 77, 9999, 5, 42)] + ["steppy-0.25-1-8", "flat"]
    m = oracle.config_model(3, final_t=2.0)  # worm-8x2, PhaseSin
    g = np.random.default_rng(4).uniform(-1, 1, (len(names), m.genome_len()))
    terrains = [oracle.create_terrain(n) for n in names]
    for prec in (32, 64):
        ctx = make_ctx(F, m, prec)
        fit, res = ctx.eval_terrains(g, terrains, want_results=True)
        assert np.isfinite(fit).all() and (res["contact_overflow"] == 0).all()
        for i, (xs, ys) in enumerate(terrains):
            mi = oracle.config_model(3, final_t=2.0)
            mi.tx, mi.ty = np.ascontiguousarray(xs), np.ascontiguousarray(ys)
            ci = make_ctx(F, mi, prec)
            alone = ci.eval(g[i:i + 1])[0]
            assert alone == fit[i], (prec, names[i], alone, fit[i])
            if prec == 64:  # the same context tracks the oracle on that terrain (placement included) before the first divergence
                _, otv, _, _ = mi.run(g[i], trace_steps=40)
                _, tv, _, _ = ci.trace(g[i], 40)
                # placement + free fall (the robot lands after ~27 steps) to rounding.  Landing on a terrain VERTEX is a
                # knife edge of the narrowphase (reference-face choice, 1- vs 2-point manifold): tests/tools/parity_probe4.py shows
                # rounding differences of 1e-13 growing ~1000x per step there, so after the landing the bound is loose
                assert np.abs(tv[:25] - otv[:25]).max() < 1e-10, (names[i], float(np.abs(tv[:25] - otv[:25]).max()))
                bounds.check(f"rnd_terrain.{names[i]}.fp64_vertex_err_40steps", np.abs(tv - otv).max(), 5e-2)
            ci.close()
        assert len(set(np.round(fit, 12))) == len(names)  # different terrains, different outcomes
        ctx.close()
    # ragged / bad input
    ctx = make_ctx(F, m, 32)
    assert ctx.eval_terrains(np.zeros((0, m.genome_len())), []).shape == (0,)
    with pytest.raises(F.VsrIllegalArgument):
        ctx.eval_terrains(g[:2], terrains[:1])
    with pytest.raises(F.VsrIllegalArgument):
        ctx.eval_terrains(g[:1], [(np.array([0.0, 0.0, 1.0]), np.array([0.0, 0.0, 0.0]))])
    ctx.close()


@pytest.mark.parametrize("shape,ctrl,sensors,terrain", [
    ("box-1x1", "phase", "", "flat"),                       # 1 voxel: 32 robots per warp, no welds
    ("box-8x4", "phase", "", "hilly-0.5-5-3"),              # 32 voxels: the largest body, one robot per warp
    ("tripod-5x3", "central", "uniform-r+vxy+t-0", "uphill-10"),  # angle + velocity + touch sensors, sloped terrain
    ("box-3x5", "dist", "uniform-a+t-0", "downhill-30"),    # tall body, distributed controller, validation terrain (Starter.java:171)
    ("comb-7x2", "hetero", "uniform-t+vxy+a-0", "flat"),    # brainHeteroDist<funsGrid: one MLP per voxel (BrainHeteroDistributed.java:76-83)
    ("biped-4x3", "hetero", "uniform-a-0", "hilly-1-10-0"),
    ("biped-4x3", "timesin", "", "flat"),                   # TimeFunctions sinusoids (BodyBrainSinusoidal.java:87-119)
    ("worm-8x2", "timesin", "", "hilly-1-10-0"),
])
def test_other_bodies_sensors_and_terrains_track_the_oracle(F, oracle, shape, ctrl, sensors, terrain):
    """Edge shapes (smallest / largest body), every sensor kind and the remaining terrain families: FP64 kernel vs oracle on the
    first 30 steps (1e-9), FP32 kernel finite and close on the first 10 steps; ragged batch sizes around the warp packing."""
    kind = {"phase": oracle.CTRL_PHASE_SIN, "central": oracle.CTRL_MLP_CENTRALIZED, "dist": oracle.CTRL_MLP_DISTRIBUTED_HOMO,
            "hetero": oracle.CTRL_MLP_DISTRIBUTED_HETERO, "timesin": oracle.CTRL_TIME_SIN}[ctrl]
    m = oracle.Model(shape, terrain, sensors, kind, oracle.default_params(1.5, 0.5))
    G = m.genome_len()
    g = np.random.default_rng(11).uniform(-1, 1, (70, G))
    if kind == oracle.CTRL_TIME_SIN:
        # the decode of bodyBrainSin (BodyBrainSinusoidal.java:96-113): frequency in [0.5, 2], phase in [0, pi], amplitude in [0, 1]
        g = g.reshape(70, -1, 3)
        g = np.stack([0.5 + 1.5 * (g[..., 0] + 1) / 2, np.pi * (g[..., 1] + 1) / 2, (g[..., 2] + 1) / 2], axis=-1).reshape(70, G)
    elif kind != oracle.CTRL_PHASE_SIN:
        # small weights keep the MLP outputs away from +-1: a fully contracted voxel sits exactly on the contraction-limit
        # knife edge (rope active or not), which flips on 1e-13 differences within a few steps (tests/tools/parity_probe5.py)
        g *= 0.15
    c64, c32 = make_ctx(F, m, 64), make_ctx(F, m, 32)
    assert c64.genome_len == G
    for r in range(3):
        _, otv, otb, otc = m.run(g[r], trace_steps=30)
        _, tv, tb, tc = c64.trace(g[r], 30)
        assert np.abs(tb - otb).max() < 1e-9 and np.abs(tc - otc).max() < 1e-9, (shape, r)
        _, tv32, _, _ = c32.trace(g[r], 10)
        assert np.abs(tv32 - otv[:10]).max() / 3.0 < 1e-4
    fit, res = c32.eval(g, want_results=True)
    assert np.isfinite(fit).all() and (res["contact_overflow"] == 0).all(), (shape, int((res["contact_overflow"] > 0).sum()))
    rpw = 32 // int(m.mask.sum())
    for n in (1, rpw, rpw + 1, 10 * rpw - 1, 10 * rpw + 1):  # around one warp / one CTA of robots
        n = min(n, len(g))
        assert np.array_equal(c32.eval(g[:n]), fit[:n]), (shape, n)
    c64.close(); c32.close()


def test_non_finite_genomes_do_not_hang_or_leak(F, oracle):
    """NaN / inf genes: the episode still terminates, the fitness of that robot is returned as is (NaN propagates like in the
    reference, whose comparator decides), and robots sharing its warp are unaffected."""
    m = oracle.config_model(1, final_t=2.0)
    ctx = make_ctx(F, m)
    g = np.random.default_rng(3).uniform(-1, 1, (9, ctx.genome_len))
    clean = ctx.eval(g)
    bad = g.copy()
    bad[1, :] = np.nan
    bad[4, 3] = np.inf
    bad[7, 10] = -np.inf
    out = ctx.eval(bad)
    ok = [0, 2, 3, 5, 6, 8]
    assert np.array_equal(out[ok], clean[ok])
    assert np.isfinite(out[[4, 7]]).all()  # tanh saturates an infinite pre-activation
    ctx.close()


def test_generation_loop_on_a_multi_device_context(F, oracle):
    """vsr_ga with a context that shards over 2 devices inside one process (peer copies) == the single-device run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    m = oracle.config_model(1, final_t=2.0)
    pops = []
    for devs in ([0], [0, 1]):
        ctx = make_ctx(F, m, devices=devs)
        ga = F.GeneticAlgorithm(ctx, 101, seed=6)
        ga.init()
        for _ in range(3):
            ga.step()
        pops.append(ga.population())
        ga.close(); ctx.close()
    assert np.array_equal(pops[0][0], pops[1][0]) and np.array_equal(pops[0][1], pops[1][1])


def test_new_controller_kinds_reduce_to_the_old_ones_on_the_device(F, oracle):
    """Same reductions as tests/test_oracle_known_answers.py, through the C ABI on the FP32 kernel: per-voxel weights that are all
    equal == the homogeneous brain, TimeFunctions sinusoids with one frequency / amplitude == PhaseSin.  Bit-identical."""
    p = oracle.default_params(3.0, 0.5)
    homo = oracle.Model("comb-7x2", "steppy-0.25-1-0", "uniform-t+vxy+a-0", oracle.CTRL_MLP_DISTRIBUTED_HOMO, p)
    het = oracle.Model("comb-7x2", "steppy-0.25-1-0", "uniform-t+vxy+a-0", oracle.CTRL_MLP_DISTRIBUTED_HETERO, p)
    g = np.random.default_rng(0).uniform(-1, 1, (37, 75))
    c1, c2 = make_ctx(F, homo), make_ctx(F, het)
    assert c2.genome_len == 825
    a, b = c1.eval(g), c2.eval(np.tile(g, (1, 11)))
    assert np.isfinite(a).all() and np.array_equal(a, b)
    g2 = np.tile(g, (1, 11))
    g2[:, 5 * 75:6 * 75] *= -1
    assert not np.array_equal(c2.eval(g2), a)
    c1.close(); c2.close()
    ph = oracle.Model("worm-8x2", "hilly-1-10-0", "", oracle.CTRL_PHASE_SIN, p, freq=1.3, amp=0.8)
    ts = oracle.Model("worm-8x2", "hilly-1-10-0", "", oracle.CTRL_TIME_SIN, p)
    phases = np.random.default_rng(1).uniform(-1, 1, (41, 16))
    triples = np.stack([np.full((41, 16), 1.3), phases, np.full((41, 16), 0.8)], axis=-1).reshape(41, 48)
    c1, c2 = make_ctx(F, ph), make_ctx(F, ts)
    assert c2.genome_len == 48
    assert np.array_equal(c1.eval(phases), c2.eval(triples))
    c1.close(); c2.close()


def test_host_transforming_mappers_through_the_population_evaluator(F, oracle, bounds):
    """brainPhaseFun<mlp (host evaluates the phase function once per voxel, device runs PhaseSin) and bodyBrainSin<dirNumGrid
    (body from the genome: robots grouped by morphology, one launch per distinct body) against the oracle run on the robots the
    mappers build."""
    from vsrevolution_b200 import locomotion as L
    ev = L.population_evaluator("biped-4x3", "", "brainPhaseFun;f=1<mlp;r=0.65;nIL=1", "flat", 3.0, 0.5)
    assert not ev.direct and ev.genome_len == 9
    g = np.random.default_rng(2).uniform(-1, 1, (12, 9))
    fit = ev.evaluate(g)
    m = oracle.Model("biped-4x3", "flat", "", oracle.CTRL_PHASE_SIN, oracle.default_params(3.0, 0.5), freq=1.0, amp=1.0)
    phases = np.stack([L.Locomotion.genome_of(ev._build(row)) for row in g])
    ref = m.eval_batch(phases)
    assert np.isfinite(fit).all()
    bounds.check("phasefun.median_abs", np.median(np.abs(fit - ref)), 2e-3)
    ev2 = L.population_evaluator("biped-4x3", "", "bodyBrainSin;p=50<dirNumGrid", "flat", 3.0, 0.5)
    assert not ev2.direct and ev2.genome_len == 36
    g2 = np.random.default_rng(3).uniform(-1, 1, (10, 36))
    fit2 = ev2.evaluate(g2)
    assert np.isfinite(fit2).all() and len(ev2.task._ctxs) >= 2   # several morphologies -> several contexts
    # every robot evaluated alone gives the same value as inside its group
    for i in (0, 4, 9):
        assert ev2.task.apply(ev2._build(g2[i])).get_velocity() == fit2[i]
    direct = L.population_evaluator("comb-7x2", "uniform-t+vxy+a-0", "brainHeteroDist;s=1<funsGrid", "flat", 2.0, 0.5)
    assert direct.direct and direct.genome_len == 825


def test_live_contexts_sharing_a_kernel_instantiation(F, oracle):
    """Several contexts alive at once (one per morphology) launch the same kernel function with different shared-memory footprints:
    creating the smaller one must not shrink the function's dynamic shared-memory opt-in under the larger one."""
    p = oracle.default_params(1.0, 0.5)
    big = oracle.Model("worm-8x2", "flat", "", oracle.CTRL_PHASE_SIN, p)
    small = oracle.Model("box-1x1", "flat", "", oracle.CTRL_PHASE_SIN, p)
    cb = make_ctx(F, big)
    gb = np.random.default_rng(0).uniform(-1, 1, (5, 16))
    first = cb.eval(gb)
    cs = make_ctx(F, small)
    assert np.isfinite(cs.eval(np.zeros((3, 1)))).all()
    assert np.array_equal(cb.eval(gb), first)
    cs.close(); cb.close()


def test_centre_trajectories_match_the_oracle_trace(F, oracle):
    """vsr_eval_centers: the (x, y) of the robot centre at every kept observation == the centre of the oracle's traced polygons
    (FP64 kernel 1e-9 while both follow the same trajectory, FP32 1e-4 on the first steps); first / last entries are the
    first_x / last_x of vsr_result; the fitness is the one vsr_eval returns."""
    m = oracle.config_model(3, final_t=2.0)
    m.params.transient_t = 0.5
    g = np.random.default_rng(1).uniform(-1, 1, (7, m.genome_len()))
    n_steps, k_first, k_last = F.episode_steps(2.0, 0.5)
    for prec in (64, 32):
        ctx = make_ctx(F, m, prec)
        fit, res, cen = ctx.eval_centers(g)
        assert cen.shape == (7, k_last - k_first + 1, 2)
        assert np.array_equal(fit, ctx.eval(g))
        assert np.array_equal(cen[:, 0, 0], res["first_x"]) and np.array_equal(cen[:, -1, 0], res["last_x"])
        if prec == 64:
            for r in (0, 3, 6):
                _, tv, _, _ = m.run(g[r], trace_steps=k_first + 25)
                ref = tv.mean(axis=2).mean(axis=1)[k_first - 1:]    # [steps][2]: mean over voxels of the mean of the 4 vertices
                assert np.abs(cen[r, :len(ref)] - ref).max() < 1e-9, (r, np.abs(cen[r, :len(ref)] - ref).max())
        ctx.close()
    # FP32 kernel: observations kept from the first step (transient 0), compared while the trajectories still coincide
    m0 = oracle.config_model(3, final_t=1.0)
    m0.params.transient_t = 0.0
    c32 = make_ctx(F, m0, 32)
    _, _, cen = c32.eval_centers(g)
    for r in (0, 3, 6):
        _, tv, _, _ = m0.run(g[r], trace_steps=10)
        assert np.abs(cen[r, :10] - tv.mean(axis=2).mean(axis=1)).max() / 3.0 < 1e-4
    c32.close()
    from vsrevolution_b200 import locomotion as L
    ev = L.population_evaluator("worm-8x2", "", "brainPhaseVals;f=1", "hilly-1-10-0", 2.0, 0.5)
    outs = ev.task.apply_all_observed([ev._build(row) for row in g[:3]])
    assert outs[1].centers.shape == (k_last - k_first + 1, 3) and len(outs[1].get_center_x_velocity_spectrum(0.0, 4.0, 8)) == 8


def test_contact_pool_overflow_is_never_silent(F, oracle, monkeypatch):
    """vsr_contact_overflows counts the warps whose contact pool overflowed in ANY launch of the context (fitness-only and generation-
    loop launches have no vsr_result to carry the flag): 0 for the shipped pool sizes, > 0 when the pool is forced too small."""
    m = oracle.config_model(1, final_t=1.5)
    g = np.random.default_rng(0).uniform(-1, 1, (64, m.genome_len()))
    ctx = make_ctx(F, m)
    ctx.eval(g)
    assert ctx.contact_overflows() == 0
    ctx.close()
    monkeypatch.setenv("VSR_POOL_SLOTS", "8")   # 3 bipeds on 8 slots: every warp overflows as soon as the robots land
    small = make_ctx(F, m)
    monkeypatch.delenv("VSR_POOL_SLOTS")
    fit, res = small.eval(g, want_results=True)
    n_warps = (64 + 2) // 3
    assert 0 < small.contact_overflows() <= n_warps and (res["contact_overflow"] > 0).any()
    # a robot that lost ground contacts never returns a rankable fitness: NaN (ranks last), finite diagnostics in vsr_result
    bad = res["contact_overflow"] > 0
    assert np.isnan(fit[bad]).all() and np.isfinite(fit[~bad]).all() and np.isfinite(res["velocity"]).all()
    fit2 = small.eval(g)                          # fitness-only launch: still counted, still NaN
    assert np.array_equal(np.isnan(fit2), bad)
    assert small.contact_overflows() > n_warps // 2
    small.close()


@pytest.mark.parametrize("cfg,n", [(2, 4000), (3, 2500), (4, 2500)])
def test_launch_shape_does_not_change_the_result(F, oracle, cfg, n):
    """Sub-wave batches run the 255-register build spread over every SM, full waves the 10-warp build (launch plan): a
    robot's fitness is bit-identical in a one-robot call, a sub-wave batch and a full-wave batch."""
    m = oracle.config_model(cfg, final_t=1.5)
    g = np.random.default_rng(cfg).uniform(-1, 1, (n, m.genome_len()))
    ctx = make_ctx(F, m)
    whole = ctx.eval(g)
    assert np.isfinite(whole).all() and ctx.contact_overflows() == 0
    assert np.array_equal(ctx.eval(g[:300]), whole[:300])
    for i in (0, 7, n - 1):
        assert ctx.eval(g[i:i + 1])[0] == whole[i]
    ctx.close()


@pytest.mark.parametrize("prec", [32, 64])
def test_kernel_sliding_on_a_slope_is_coulomb_plus_damping(F, oracle, prec):
    """The closed-form check of tests/test_oracle_known_answers.py applied to the CUDA kernels themselves (no oracle in the loop): a 1x1
    robot with mu = 0.1 on the `downhill-30` incline slides with v' = (v + g (sin theta - mu cos theta) dt)(1 - dt d)."""
    import math
    tx, ty = oracle.create_terrain("downhill-30")
    e = F.env_default()
    e.final_t, e.transient_t, e.precision, e.voxel_friction = 5.0, 0.5, prec, 0.05
    ctx = F.Context(e, tx, ty)
    ctx.body_set(np.ones((1, 1), dtype=np.uint8), [])
    ctx.controller_set(F.CTRL_PHASE_SIN, (), 0, 1.0, 0.0)
    _, _, tb, _ = ctx.trace(np.zeros(1), 300)
    th = math.atan2(ty[1] - ty[2], tx[2] - tx[1])
    vel = tb[:, :, 3:5].mean(axis=1)
    along = vel[:, 0] * math.cos(th) - vel[:, 1] * math.sin(th)
    a, dt, k0 = 9.8 * (math.sin(th) - math.sqrt(0.05 * 0.2) * math.cos(th)), 1 / 60.0, 120
    v, rec = along[k0], []
    for _k in range(k0, 300):
        rec.append(v)
        v = (v + a * dt) * (1 - dt)
    err = np.abs(along[k0:] - np.array(rec)).max() / abs(rec[-1])
    assert err < (1e-6 if prec == 64 else 2e-4), err
    ctx.close()


# ---- robots that actually locomote: contacts and the solver's knife-edge flips dominate here -------------
def _cmp(bounds, tag, fit, ref, default_rho, default_med):
    d = np.abs(fit - ref)
    rel = d / np.maximum(np.abs(ref), 1e-9)
    rho = spearman(fit, ref)
    print(f"{tag}: n={len(ref)} mean|v|={np.abs(ref).mean():.3f} spearman={rho:.4f} median|d|={np.median(d):.4g} "
          f"median rel={np.median(rel):.4g} p90 rel={np.quantile(rel, 0.9):.4g}")
    bounds.check(tag + ".spearman", rho, default_rho, kind="spearman")
    bounds.check(tag + ".median_rel", np.median(rel), default_med)
    bounds.check(tag + ".p90_rel", np.quantile(rel, 0.9), 10 * default_med)


def test_evolved_population_parity(F, oracle, bounds):
    """Config 2 after 20 generations of the device numGA loop (Starter.java:174,427-435: what the solver actually evaluates): the
    evolved bipeds run at several units/s; FP32 and FP64 kernels against the oracle, with the FP64 kernel (same algorithm, other
    rounding order) as the chaotic floor the FP32 build has to stay at."""
    from vsrevolution_b200.locomotion import population_evaluator
    T, n = 15.0, 2048
    ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", T, 1.0)
    ga = F.GeneticAlgorithm(ev.ctx, n, seed=0)
    ga.init()
    for _ in range(30):
        ga.step()
    g, f = ga.population()
    ga.close()
    assert ev.ctx.contact_overflows() == 0
    ev.ctx.close()
    m = oracle.config_model(2, final_t=T)
    ref = m.eval_batch(g, n_threads=os.cpu_count() or 8)
    print(f"evolved population: median oracle velocity {np.median(ref):.3f}, best {ref.max():.3f}")
    assert np.median(np.abs(ref)) > 0.3, "the evolved population must locomote for this test to mean anything"
    fits = {}
    for prec in (32, 64):
        ctx = make_ctx(F, m, prec)
        fits[prec] = ctx.eval(g)
        assert np.isfinite(fits[prec]).all() and ctx.contact_overflows() == 0
        ctx.close()
    assert np.array_equal(fits[32], f)                       # the loop's own fitness values are those of a fresh evaluation
    _cmp(bounds, "evolved2.fp32_vs_oracle", fits[32], ref, 0.8, 0.2)
    _cmp(bounds, "evolved2.fp64_vs_oracle", fits[64], ref, 0.8, 0.2)
    # FP32 must sit at the chaotic floor: not worse than 1.5x the FP64 kernel's own distance from the oracle
    d32, d64 = np.median(np.abs(fits[32] - ref)), np.median(np.abs(fits[64] - ref))
    bounds.check("evolved2.fp32_over_fp64_median_ratio", d32 / d64, 3.0)


def test_moving_worms_parity(F, oracle, bounds):
    """Config 3 robots filtered to those that move (|v| > 0.3 in the oracle): PhaseSin worms on hilly terrain."""
    T, n = 20.0, 2048
    m = oracle.config_model(3, final_t=T)
    g = np.random.default_rng(1).uniform(-1, 1, (n, m.genome_len()))
    ref = m.eval_batch(g, n_threads=os.cpu_count() or 8)
    mv = np.abs(ref) > 0.3
    assert mv.sum() >= 20, int(mv.sum())
    fits = {}
    for prec in (32, 64):
        ctx = make_ctx(F, m, prec)
        fits[prec] = ctx.eval(g)
        assert np.isfinite(fits[prec]).all() and ctx.contact_overflows() == 0
        ctx.close()
    _cmp(bounds, "worms3.all.fp32_vs_oracle", fits[32], ref, 0.5, 2.0)
    _cmp(bounds, "worms3.all.fp64_vs_oracle", fits[64], ref, 0.5, 2.0)
    _cmp(bounds, "worms3.movers.fp32_vs_oracle", fits[32][mv], ref[mv], 0.3, 1.0)
    _cmp(bounds, "worms3.movers.fp64_vs_oracle", fits[64][mv], ref[mv], 0.3, 1.0)


def test_config4_population_parity(F, oracle, bounds):
    """Config 4 (comb-7x2, DistributedSensing, touch / velocity / area sensors, steppy terrain) on 1024 robots x 10 s."""
    T, n = 10.0, 1024
    m = oracle.config_model(4, final_t=T)
    g = np.random.default_rng(2).uniform(-1, 1, (n, m.genome_len()))
    ref = m.eval_batch(g, n_threads=os.cpu_count() or 8)
    for prec in (32, 64):
        ctx = make_ctx(F, m, prec)
        fit = ctx.eval(g)
        assert np.isfinite(fit).all() and ctx.contact_overflows() == 0
        ctx.close()
        d = np.abs(fit - ref)
        bounds.check(f"pop4.fp{prec}.spearman", spearman(fit, ref), 0.8, kind="spearman")
        bounds.check(f"pop4.fp{prec}.median_abs", np.median(d), 2e-3)
        bounds.check(f"pop4.fp{prec}.p90_abs", np.quantile(d, 0.9), 2e-2)


# ---- vsr_eval_ex: robot transformations and the per-observation summaries behind Outcome's other getters ------------------------
@pytest.mark.parametrize("cfg", [1, 3, 4])
def test_observation_summaries_match_the_oracle(F, oracle, bounds, cfg):
    """Average posture (Starter.java:180-190), footprints (NamedFunctions.java:179-180,392) and the centre + central-voxel-angle
    trajectory (NamedFunctions.java:139-172) from vsr_eval_ex against the oracle's vo_run_ex on windows that precede the first
    discrete divergence (42 steps for the FP64 kernel; the first 10 observations for the FP32 kernel): FP64 kernel exact (posture,
    footprint bytes) / 1e-9 (centres); FP32 kernel: same footprints and centres within 1e-4 of the voxel side over its window, at
    most 2 posture cells different over the 42 steps."""
    g = None
    for prec, transient in ((64, 0.2), (32, 0.0)):
        m = oracle.config_model(cfg, final_t=0.7)
        m.params.transient_t = transient
        if g is None:
            g = np.random.default_rng(CFG_SEED[cfg]).uniform(-1, 1, (7, m.genome_len()))
        refs = [m.run_ex(row, posture_n=8, footprint_n=(4, 3), centers=True) for row in g]
        K = refs[0][1]["kept"]
        ctx = make_ctx(F, m, prec)
        assert ctx.central_voxel == m.central_voxel()
        fit, res, ex = ctx.eval_ex(g, posture_n=8, footprint_n=(4, 3), centers=True)
        assert ex["kept"] == K and ex["posture"].shape == (7, 64) and ex["centers"].shape == (7, K, 3)
        plain = ctx.eval(g)
        bounds.check(f"obs.cfg{cfg}.fp{prec}.fitness_vs_plain_kernel", np.abs(fit - plain).max(), 1e-12 if prec == 64 else 1e-3)
        for r in range(7):
            o = refs[r][1]
            dpost = int((ex["posture"][r] != o["posture"]).sum())
            if prec == 64:
                dfp = [int((ex["footprints"][q][r] != o["footprints"][q]).sum()) for q in range(2)]
                assert dpost == 0 and dfp == [0, 0], (cfg, r, dpost, dfp)
                assert np.abs(ex["centers"][r] - o["centers"]).max() < 1e-9
                assert abs(fit[r] - refs[r][0].velocity) < 1e-9
            else:
                dfp = [int((ex["footprints"][q][r][:10] != o["footprints"][q][:10]).sum()) for q in range(2)]
                assert dpost <= 2 and dfp == [0, 0], (cfg, r, dpost, dfp)
                bounds.check(f"obs.cfg{cfg}.fp32.center_err_first10", np.abs(ex["centers"][r][:10, :2] - o["centers"][:10, :2]).max() / 3.0, 1e-4)
                bounds.check(f"obs.cfg{cfg}.fp32.angle_err_first10", np.abs(ex["centers"][r][:10, 2] - o["centers"][:10, 2]).max(), 1e-4)
        assert ex["posture"].any()
        ctx.close()
    # only some requests; none; bad requests
    m = oracle.config_model(cfg, final_t=0.7)
    m.params.transient_t = 0.2
    K = oracle.episode_steps(0.7, 0.2)[2] - oracle.episode_steps(0.7, 0.2)[1] + 1
    ctx = make_ctx(F, m, 32)
    f2, _, ex2 = ctx.eval_ex(g, footprint_n=(0, 5))
    assert ex2["posture"] is None and ex2["footprints"][0] is None and ex2["footprints"][1].shape == (7, K) and (ex2["footprints"][1] < 32).all()
    assert np.array_equal(ctx.eval_ex(g)[0], ctx.eval(g))
    with pytest.raises(F.VsrError):
        ctx.eval_ex(g, posture_n=9)
    ctx.close()


def test_broken_transformation_freezes_actuators(F, oracle, bounds):
    """`broken-<ratio>-<seed>` (RobotUtils.buildRobotTransformation, Starter.java:126,428): the masks replay java.util.Random, a frozen
    voxel applies no force (control energy of an all-broken robot is 0), the FP64 kernel equals the oracle with the same mask, and a
    batch with masks is bit-identical to each robot evaluated alone."""
    # 24 steps: the robots are still falling (they land after ~27 steps), so no contact knife edge can separate the FP64 kernel from
    # the oracle, while the control energy (sum f^2 dt over the voxels) already tells frozen voxels from working ones
    m = oracle.config_model(1, final_t=0.4)
    m.params.transient_t = 0.1
    g = np.random.default_rng(3).uniform(-1, 1, (9, m.genome_len()))
    ctx = make_ctx(F, m, 64)
    masks = ctx.broken_masks(0.4, 11, 9)
    # independent replay of `random.nextDouble() > ratio ? v : new BreakableVoxel(..., random.nextInt(), ...)`
    jr = oracle.JavaRandom(11)
    exp = []
    for _ in range(9):
        mk = 0
        for v in range(10):
            if not (jr.next_double() > 0.4):
                mk |= 1 << v
                jr.next_int()
        exp.append(mk)
    assert masks.tolist() == exp and 0 < sum(bin(x).count("1") for x in exp) < 90
    st = F.JavaRandomState(11)
    assert np.array_equal(np.concatenate([ctx.broken_masks_stream(0.4, st, 4), ctx.broken_masks_stream(0.4, st, 5)]), masks)
    fit, res, _ = ctx.eval_ex(g, frozen_masks=masks)
    unfrozen = ctx.eval(g, want_results=True)[1]
    for r in range(9):
        o, _ = m.run_ex(g[r], frozen_mask=int(masks[r]))
        assert abs(fit[r] - o.velocity) < 1e-9, (r, fit[r], o.velocity)
        for name in ("control_energy", "area_ratio_abs", "area_ratio_sq"):
            assert abs(res[name][r] - getattr(o, name)) < 1e-9 * max(1.0, getattr(o, name)), (r, name)
        if masks[r]:
            assert res["control_energy"][r] < unfrozen["control_energy"][r]
    allb = np.full(9, (1 << 10) - 1, dtype=np.uint32)
    _, res_all, _ = ctx.eval_ex(g, frozen_masks=allb)
    assert (res_all["control_energy"] == 0).all() and len(set(np.round(res_all["velocity"], 12))) == 1
    assert not np.array_equal(ctx.eval_ex(g, frozen_masks=np.zeros(9, dtype=np.uint32))[0], fit) or not masks.any()
    assert np.array_equal(ctx.eval_ex(g, frozen_masks=np.zeros(9, dtype=np.uint32))[0], ctx.eval(g))
    with pytest.raises(F.VsrIllegalArgument):
        ctx.eval_ex(g, frozen_masks=np.full(9, 1 << 10, dtype=np.uint32))
    ctx.close()
    c32 = make_ctx(F, m, 32)
    whole = c32.eval_ex(g, frozen_masks=masks)[0]
    for r in (0, 4, 8):
        assert c32.eval_ex(g[r:r + 1], frozen_masks=masks[r:r + 1])[0][0] == whole[r]
    # staged for the device-resident path
    import torch
    dg = torch.from_numpy(g).cuda()
    dfit = torch.zeros(9, dtype=torch.float64, device="cuda")
    c32.stage_eval_opts(9, frozen_masks=masks)
    c32.eval_device(dg.data_ptr(), 9, dfit.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(dfit.cpu().numpy(), whole)
    c32.stage_eval_opts(0)
    c32.eval_device(dg.data_ptr(), 9, dfit.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(dfit.cpu().numpy(), c32.eval(g))
    c32.close()
    # host mirror: validate with a transformation, Outcome getters of the observed path
    from vsrevolution_b200.builder import build_mapper
    from vsrevolution_b200.locomotion import build_locomotion_task, validate
    from vsrevolution_b200.robot import Robot, RobotUtils
    target = Robot(None, RobotUtils.build_sensorizing_function("uniform-a-0")(RobotUtils.build_shape("biped-4x3")))
    mapper = build_mapper("brainCentralized<mlp;r=0.65;nIL=1").build_for(target)
    robots = [mapper(row) for row in g[:5]]
    vb = validate(robots, "flat", "broken-0.4-11", 0, 2.0, 0.5)
    vi = validate(robots, "flat", "identity", 0, 2.0, 0.5)
    assert [v.outcome.get_velocity() for v in vb] != [v.outcome.get_velocity() for v in vi]
    tr = RobotUtils.build_robot_transformation("broken-0.4-11")
    assert [RobotUtils.frozen_mask(tr(r)) for r in robots] == exp[:5]
    with pytest.raises(F.VsrError):
        validate(robots, "flat", "breakable-time-10/1-5/1-0", 0, 2.0, 0.5)
    outs = build_locomotion_task("flat", 2.0, 0.5).apply_all_observed(robots)
    o0 = outs[0]
    assert o0.get_average_posture(8).shape == (8, 8) and o0.get_average_posture(8).any()
    assert len(o0.get_footprints_spectra(4, 0.0, 5.0, 8)) == 4 and len(o0.get_footprints_spectra(3, 0.0, 5.0, 4)[0]) == 4
    assert len(o0.get_center_angle_spectrum(0.0, 5.0, 8)) == 8 and np.isfinite(o0.get_center_x_velocity_spectrum(0.0, 5.0, 8)).all()


def test_broken_robots_population_parity(F, oracle, bounds):
    """Full-length statistical parity with the `broken-0.3` transformation applied: FP32 kernel vs the oracle run with the same masks."""
    T, n = 5.0, 256
    m = oracle.config_model(2, final_t=T)
    g = np.random.default_rng(8).uniform(-1, 1, (n, m.genome_len()))
    ctx = make_ctx(F, m, 32)
    masks = ctx.broken_masks(0.3, 5, n)
    fit = ctx.eval_ex(g, frozen_masks=masks)[0]
    ctx.close()
    ref = np.array([m.run_ex(g[r], frozen_mask=int(masks[r]))[0].velocity for r in range(n)])
    d = np.abs(fit - ref)
    bounds.check("broken2.fp32.spearman", spearman(fit, ref), 0.9, kind="spearman")
    bounds.check("broken2.fp32.median_abs", np.median(d), 1e-3)


@pytest.mark.parametrize("shape,sensors,terrain,ctrl", [
    ("biped-4x3", "spinedTouchSighted-t-f-0", "hilly-1-10-0", "central"),
    ("biped-4x3", "uniform-px+py+l5+a-0", "flat", "central"),
    ("worm-6x2", "spinedTouch-t-f-0.01", "steppy-0.25-1-0", "central"),
    ("comb-7x2", "uniform-py+lf5+t-0", "hilly-1-10-0", "dist"),
])
def test_per_voxel_sensor_lists_and_the_extra_sensor_vocabulary(F, oracle, bounds, shape, sensors, terrain, ctrl):
    """RobotUtils.buildSensorizingFunction names whose lists differ between voxels (`spinedTouch[Sighted]`, and `uniform-` over the
    position-dependent predefined sensors px / py / l5) with the Constant, TimeFunction (cpg), Lidar and hard-Normalization sensors:
    the library's own name layer against the oracle's (bit-exact), FP64 kernel == oracle over 40 steps, FP32 kernel tracks for 10."""
    kind = oracle.CTRL_MLP_CENTRALIZED if ctrl == "central" else oracle.CTRL_MLP_DISTRIBUTED_HOMO
    m = oracle.Model(shape, terrain, sensors, kind, oracle.default_params(2.0, 0.5), signals=1)
    # name layer: vsr_sensor_config_voxels == oracle/model.py sensor_config_voxels
    lib_s, lib_off = F.sensor_config_voxels(sensors, m.mask)
    assert lib_off.tolist() == m.vox_off and len(lib_s) == len(m.vox_sensors)
    for a, b in zip(lib_s, m.vox_sensors):
        ka = a.key()
        kb = (b.kind, b.axes, b.rotated, b.soft_norm, b.dom_min, b.dom_max, b.avg_window, b.noise_sigma, b.param, b.n_rays, b.side)
        assert ka == kb, (ka, kb)
    # small weights keep the (large) MLPs' outputs away from +-1: a fully contracted voxel sits exactly on the contraction-limit knife
    # edge (rope lower limit == spring minimum), where 1e-13 rounding differences flip a rope within a few steps even in free fall
    g = 0.15 * np.random.default_rng(12).uniform(-1, 1, (3, m.genome_len()))
    c64 = make_ctx(F, m, 64)
    assert c64.genome_len == m.genome_len()
    cell, sensor, reading = c64.input_map()
    if ctrl == "central":
        assert len(cell) == m.ctrl.layers[0]
    for r in range(3):
        _, otv, otb, otc = m.run(g[r], trace_steps=40)
        _, tv, tb, tc = c64.trace(g[r], 40)
        # free fall (the robots land after ~27 steps) to rounding: every sensor but Touch is already exercised there; after the landing
        # rounding differences grow ~1000x per step at contact knife edges (see test_per_evaluation_terrains...), so the bound is looser
        assert np.abs(tb[:25] - otb[:25]).max() < 1e-10 and np.abs(tc[:25] - otc[:25]).max() < 1e-10, (r, np.abs(tb[:25] - otb[:25]).max())
        bounds.check(f"voxsens.{shape}.{sensors}.fp64.body_err_40steps", np.abs(tb - otb).max(), 0.5)
    c64.close()
    c32 = make_ctx(F, m, 32)
    for r in range(3):
        _, otv, _, otc = m.run(g[r], trace_steps=10)
        _, tv, _, tc = c32.trace(g[r], 10)
        bounds.check(f"voxsens.{shape}.{sensors}.fp32.vertex_err_first10", np.abs(tv - otv).max() / 3.0, 1e-4)
        bounds.check(f"voxsens.{shape}.{sensors}.fp32.ctrl_err_first10", np.abs(tc - otc).max(), 1e-3)
    whole = c32.eval(g)
    assert np.isfinite(whole).all() and c32.eval(g[1:2])[0] == whole[1]
    c32.close()
    # host mirror: the same robot through RobotUtils.build_sensorizing_function + the mapper
    from vsrevolution_b200.locomotion import population_evaluator
    mapper = "brainCentralized<mlp;r=0.65;nIL=1" if ctrl == "central" else "brainHomoDist;s=1<mlp;r=0.65;nIL=1"
    ev = population_evaluator(shape, sensors, mapper, terrain, 2.0, 0.5)
    assert ev.genome_len == m.genome_len() and np.array_equal(ev.evaluate(g), whole)


def test_observation_build_is_launch_shape_independent(F, oracle):
    """vsr_eval_ex runs the default robot on the flexible 8-warp observation build for batches of up to 8 robot-warps per SM and on the
    10-warp / 168-register one above (4000 bipeds = 9 warps per SM stay one wave).  Both carry the specialisations of the plain
    kernels, so fitness, posture, footprints and centres of a robot are bit-identical in either shape, and the fitness is the one
    vsr_eval returns."""
    m = oracle.config_model(2, final_t=1.5)
    g = np.random.default_rng(21).uniform(-1, 1, (4000, m.genome_len()))
    ctx = make_ctx(F, m, 32)
    fit, res, ex = ctx.eval_ex(g, posture_n=8, footprint_n=(4, 3), centers=True)
    assert ctx.kernel_info()["threads_per_cta"] == 320 and ctx.kernel_info()["ctas_per_sm"] == 1
    f2, r2, e2 = ctx.eval_ex(g[:300], posture_n=8, footprint_n=(4, 3), centers=True)
    assert ctx.kernel_info()["threads_per_cta"] < 320
    assert np.isfinite(fit).all() and ctx.contact_overflows() == 0
    assert np.array_equal(fit[:300], f2) and np.array_equal(ex["posture"][:300], e2["posture"])
    assert np.array_equal(ex["footprints"][0][:300], e2["footprints"][0]) and np.array_equal(ex["centers"][:300], e2["centers"])
    assert ex["posture"].any(axis=1).all()
    assert np.array_equal(ctx.eval(g), fit)
    ctx.close()


@pytest.mark.parametrize("cfg", [3, 4])
def test_height_map_broadphase_is_a_filter_only(F, oracle, cfg, monkeypatch):
    """The coarse height map of the terrain only thins out the narrowphase candidates: fitness and vsr_result of every
    robot are bit-identical with the map and with the old test against the highest terrain vertex (VSR_NO_HEIGHT_MAP)."""
    m = oracle.config_model(cfg, final_t=4.0)
    g = np.random.default_rng(40 + cfg).uniform(-1, 1, (600, m.genome_len()))
    ctx = make_ctx(F, m)
    fit, res = ctx.eval(g, want_results=True)
    ctx.close()
    monkeypatch.setenv("VSR_NO_HEIGHT_MAP", "1")
    ref = make_ctx(F, m)
    monkeypatch.delenv("VSR_NO_HEIGHT_MAP")
    fit0, res0 = ref.eval(g, want_results=True)
    ref.close()
    assert np.isfinite(fit).all() and np.array_equal(fit, fit0) and res.tobytes() == res0.tobytes()
