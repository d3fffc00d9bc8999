"""Known-answer tests that pin the CPU oracle (the reference ships no tests or golden vectors).

Each test derives the expected numbers independently (closed form / a few lines of numpy), not from the
oracle's own code.
"""
import ctypes as C
import math

import numpy as np


def test_java_random_known_values(oracle):
    # Published java.util.Random outputs.
    assert oracle.JavaRandom(0).next_int() == -1155484576
    assert oracle.JavaRandom(42).next_int() == -1170105035
    assert oracle.JavaRandom(0).next_double() == 0.730967787376657
    assert oracle.JavaRandom(0).next_gaussian() == 0.8025330637390305
    r = oracle.JavaRandom(0)
    assert [r.next_int(10) for _ in range(5)] == [0, 8, 9, 7, 5]


def test_free_fall_semi_implicit_euler(oracle):
    p = oracle.default_params()
    n = 120
    y = np.zeros(n)
    vy = np.zeros(n)
    oracle.lib().vo_test_free_fall(C.byref(p), n, y.ctypes.data_as(C.c_void_p), vy.ctypes.data_as(C.c_void_p))
    dt, g, c = 1 / 60, -9.8, 1.0
    ev, ey, v, yy = [], [], 0.0, 0.0
    for _ in range(n):
        v = (v + g * dt) * (1 - dt * c)  # dyn4j: gravity, then damping factor clamp(1 - dt*c, 0, 1)
        yy += v * dt
        ev.append(v)
        ey.append(yy)
    np.testing.assert_allclose(vy, ev, rtol=0, atol=1e-13)
    np.testing.assert_allclose(y, ey, rtol=0, atol=1e-12)
    # terminal velocity of that recurrence: v* = g*dt*(1-dt*c)/(dt*c)
    assert abs(ev[-1] - g * (1 - dt * c) / c * (1 - (1 - dt * c) ** n)) < 1e-9


def test_soft_distance_joint_is_implicit_damped_oscillator(oracle):
    p = oracle.default_params()
    p.gravity_y = 0.0
    p.lin_damping = 0.0
    p.ang_damping = 0.0
    n = 90
    out = np.zeros(n)
    rest, x0 = 2.0, 2.3
    oracle.lib().vo_test_spring_pair(C.byref(p), C.c_double(rest), C.c_double(x0), n, out.ctypes.data_as(C.c_void_p))
    # Box2D/dyn4j soft constraint == u' = (u - h w^2 x) / (1 + 2 z w h + h^2 w^2); x' = x + h u'
    h, w, z = 1 / 60, 2 * math.pi * 8.0, 0.3
    x, u, exp = x0 - rest, 0.0, []
    for _ in range(n):
        u = (u - h * w * w * x) / (1 + 2 * z * w * h + h * h * w * w)
        x += h * u
        exp.append(rest + x)
    np.testing.assert_allclose(out, exp, rtol=0, atol=1e-11)
    assert abs(out[-1] - rest) < 1e-3  # it is damped


def _drop(oracle, p, angle, height, vx, n):
    st = np.zeros((n, 6))
    oracle.lib().vo_test_box_drop(C.byref(p), C.c_double(angle), C.c_double(height), C.c_double(vx), n,
                                  st.ctypes.data_as(C.c_void_p))
    return st


def test_box_rests_on_flat_ground_within_slop(oracle):
    p = oracle.default_params()
    st = _drop(oracle, p, 0.0, 0.45 + 0.02, 0.0, 240)
    half = 0.45
    pen = half - st[-1, 1]
    assert 0 <= pen <= p.linear_tol + 1e-9, pen  # resting penetration stays inside the allowed slop
    assert abs(st[-1, 4]) < 1e-9 and abs(st[-1, 3]) < 1e-12 and abs(st[-1, 2]) < 1e-12


def test_box_restitution_above_threshold(oracle):
    p = oracle.default_params()
    p.lin_damping = 0.0
    st = _drop(oracle, p, 0.0, 0.45 + 1.0, 0.0, 60)
    vy = st[:, 4]
    k = int(np.argmax(vy > 0))  # first step moving up
    v_in = vy[k - 1] + p.gravity_y * p.dt  # velocity entering the solver of step k (after gravity)
    assert v_in < -p.restitution_velocity
    # e = max(0.1, 0) = 0.1
    assert abs(vy[k] - (-0.1 * v_in)) < 1e-9


def test_box_friction_is_coulomb(oracle):
    p = oracle.default_params()
    p.lin_damping = 0.0
    p.voxel_friction = 0.45  # mixed mu = sqrt(0.45 * 0.2) = 0.3 < 1: slides flat, no tipping
    mu = math.sqrt(0.45 * 0.2)
    st = _drop(oracle, p, 0.0, 0.45 - 0.004, 3.0, 120)
    vx = st[:, 3]
    sliding = vx[:-1] > 0.2
    # while sliding flat, the total normal impulse is m g dt, so dv = -mu g dt per step
    np.testing.assert_allclose(np.diff(vx)[sliding][3:], -mu * 9.8 / 60, rtol=0, atol=1e-6)
    assert abs(vx[-1]) < 1e-9 and abs(st[-1, 5]) < 1e-9 and np.abs(st[:, 2]).max() < 1e-6
    # stopping distance v0^2 / (2 mu g) up to one step of discretisation
    assert abs(st[-1, 0] - 9.0 / (2 * mu * 9.8)) < 3.0 / 60 + 1e-3
    # frictionless control: mu = 0 keeps vx
    p.voxel_friction = 0.0
    st0 = _drop(oracle, p, 0.0, 0.45 - 0.004, 3.0, 120)
    assert abs(st0[-1, 3] - 3.0) < 1e-12 and abs(st0[-1, 0] - 6.0) < 1e-9


def test_tilted_box_settles_flat(oracle):
    p = oracle.default_params()
    st = _drop(oracle, p, 0.3, 1.2, 0.0, 600)
    a = st[-1, 2]
    assert abs(a - round(a / (math.pi / 2)) * (math.pi / 2)) < 0.02
    assert abs(st[-1, 4]) < 1e-6


def test_robot_at_rest_lengths_has_zero_internal_force(oracle):
    """With gravity off and f=0 every spring is at its rest length and every weld is satisfied: nothing moves.
    This pins the anchor tables of Voxel.assemble() and the weld anchors of Robot.assemble()."""
    p = oracle.default_params(2.0, 0.5)
    p.gravity_y = 0.0
    for shape in ("biped-4x3", "worm-8x2", "comb-7x2"):
        m = oracle.Model(shape, "flat", "", oracle.CTRL_PHASE_SIN, p, amp=0.0)
        res, tv, tb, tc = m.run(np.zeros(m.n_voxels), trace_steps=60)
        assert np.abs(tb[:, :, 3:]).max() < 1e-12
        assert np.abs(tb[-1, :, :3] - tb[0, :, :3]).max() < 1e-12
        # geometry: voxel polygons are 3x3 squares on a 3-grid, one unit above the ground
        v0 = tv[0]
        np.testing.assert_allclose(v0[:, 1, 0] - v0[:, 0, 0], 3.0, atol=1e-12)
        np.testing.assert_allclose(v0[:, 0, 1] - v0[:, 3, 1], 3.0, atol=1e-12)
        assert abs(v0[:, :, 0].min() - 11.0) < 1e-12 and abs(v0[:, :, 1].min() - 6.0) < 1e-12


def test_spring_ranges_follow_area_ratio_delta(oracle):
    """f=+1 shrinks every spring to range.min: the voxel's steady-state area ratio approaches 1-delta (free space)."""
    p = oracle.default_params(3.0, 0.5)
    p.gravity_y = 0.0
    for sign, target in ((+1.0, 1 - 0.225), (-1.0, 1 + 0.225)):
        m = oracle.Model("box-1x1", "flat", "", oracle.CTRL_PHASE_SIN, p, amp=sign, freq=0.0)
        res, tv, tb, tc = m.run(np.array([math.pi / 2]), trace_steps=180)
        q = tv[-1, 0]
        area = 0.5 * abs(sum(q[i, 0] * q[(i + 1) % 4, 1] - q[(i + 1) % 4, 0] * q[i, 1] for i in range(4)))
        assert abs(area / 9.0 - target) < 2e-3, (area / 9.0, target)
        assert abs(tc[-1, 0] - sign) < 1e-12


def test_determinism_and_counts(oracle):
    m = oracle.config_model(1, final_t=5.0)
    g = np.random.default_rng(0).uniform(-1, 1, m.genome_len())
    r1, tv1, *_ = m.run(g, trace_steps=100)
    r2, tv2, *_ = m.run(g, trace_steps=100)
    assert r1.velocity == r2.velocity and np.array_equal(tv1, tv2)
    # every velocity iteration visits every spring and weld exactly once
    assert r1.n_spring_vel == r1.n_steps * 10 * 180
    assert r1.n_weld_vel == r1.n_steps * 10 * 24
    fit = m.eval_batch(np.stack([g, g, -g]), n_threads=3)
    assert fit[0] == r1.velocity == fit[1]


def test_noisy_sensor_adds_first_gaussians_of_random_0(oracle):
    """Noisy(sensor, sigma, 0): the c-th reading of a 1-dim sensor gets the c-th Gaussian of Random(0) times sigma.
    Checked through a 1-voxel robot whose MLP is the identity-like map out = tanh(w * tanh(reading))."""
    p = oracle.default_params(0.2, 0.0)
    p.gravity_y = 0.0
    outs = {}
    for cfg in ("uniform-a-0", "uniform-a-0.05"):
        m = oracle.Model("box-1x1", "flat", cfg, oracle.CTRL_MLP_CENTRALIZED, p, n_inner=0)
        assert m.genome_len() == 2  # bias + 1 weight
        _, _, _, tc = m.run(np.array([0.0, 1.0]), trace_steps=3)
        outs[cfg] = tc[:, 0]
    r = oracle.JavaRandom(0)
    gs = [r.next_gaussian() for _ in range(3)]
    for k in range(3):
        clean_reading = math.atanh(math.atanh(outs["uniform-a-0"][k]))
        noisy_reading = math.atanh(math.atanh(outs["uniform-a-0.05"][k]))
        # step 1 is exact; later steps feed back through the controlled area, so only the first is a closed form
        if k == 0:
            assert abs(noisy_reading - clean_reading - gs[0] * 0.05) < 1e-12


def test_new_controller_kinds_reduce_to_the_old_ones(oracle):
    """DISTRIBUTED_HETERO with the same weights in every voxel == DISTRIBUTED_HOMO (BrainHeteroDistributed.java:76-83 installs one
    function per voxel into the same DistributedSensing); TIME_SIN with one frequency / amplitude == PhaseSin
    (BodyBrainSinusoidal.java:115 vs PhaseSin's a*sin(2 pi f t + phi)).  Bit-identical: same arithmetic, same order."""
    p = oracle.default_params(2.0, 0.5)
    homo = oracle.Model("comb-7x2", "flat", "uniform-t+vxy+a-0", oracle.CTRL_MLP_DISTRIBUTED_HOMO, p)
    het = oracle.Model("comb-7x2", "flat", "uniform-t+vxy+a-0", oracle.CTRL_MLP_DISTRIBUTED_HETERO, p)
    assert homo.genome_len() == 75 and het.genome_len() == 11 * 75
    g = np.random.default_rng(0).uniform(-1, 1, (3, 75))
    assert np.array_equal(homo.eval_batch(g), het.eval_batch(np.tile(g, (1, 11))))
    # and it is really per voxel: changing one voxel's weights changes the outcome
    g2 = np.tile(g, (1, 11))
    g2[:, 5 * 75:6 * 75] *= -1
    assert not np.array_equal(het.eval_batch(g2), homo.eval_batch(g))
    ph = oracle.Model("biped-4x3", "flat", "", oracle.CTRL_PHASE_SIN, p, freq=1.3, amp=0.8)
    ts = oracle.Model("biped-4x3", "flat", "", oracle.CTRL_TIME_SIN, p)
    assert ts.genome_len() == 30
    phases = np.random.default_rng(1).uniform(-1, 1, (3, 10))
    triples = np.stack([np.full((3, 10), 1.3), phases, np.full((3, 10), 0.8)], axis=-1).reshape(3, 30)
    assert np.array_equal(ph.eval_batch(phases), ts.eval_batch(triples))


def test_sliding_on_a_slope_is_coulomb_plus_damping(oracle):
    """A 1x1 robot on `downhill-30` (Starter.java:171's validation terrain; the polyline drops 1980 sin(30 deg) over 1980, i.e. an
    incline of atan(1/2)): with mu = sqrt(f_voxel f_ground) < tan(theta) the centre of mass slides down the slope with
    v' = (v + g (sin theta - mu cos theta) dt)(1 - dt d) -- kinetic Coulomb friction on an inclined polygon plus dyn4j's linear damping
    form -- and with the default friction (mu = 4.47) it sticks.  Pins the contact normal / tangent of inclined ground polygons."""
    dt = 1 / 60.0
    for f_voxel in (0.05, 0.2):
        p = oracle.default_params(5.0, 0.5)
        p.voxel_friction = f_voxel
        m = oracle.Model("box-1x1", "downhill-30", "", oracle.CTRL_PHASE_SIN, p, amp=0.0)
        th = math.atan2(m.ty[1] - m.ty[2], m.tx[2] - m.tx[1])
        assert abs(math.tan(th) - 0.5) < 1e-12
        _, _, tb, _ = m.run(np.zeros(1), trace_steps=300)
        vel = tb[:, :, 3:5].mean(axis=1)
        along = vel[:, 0] * math.cos(th) - vel[:, 1] * math.sin(th)
        normal = vel[:, 0] * math.sin(th) + vel[:, 1] * math.cos(th)
        mu = math.sqrt(f_voxel * 0.2)
        a = 9.8 * (math.sin(th) - mu * math.cos(th))
        k0 = 120                                     # landed and settled
        v, rec = along[k0], []
        for _k in range(k0, 300):
            rec.append(v)
            v = (v + a * dt) * (1 - dt * 1.0)
        assert np.abs(along[k0:] - np.array(rec)).max() < 1e-6 * abs(rec[-1])
        assert np.abs(normal[k0:]).max() < 1e-5      # it stays on the slope
        assert abs(along[-1] - a * (1 - dt)) < 0.03 * a   # close to the terminal velocity a dt (1 - dt d) / (dt d)
    p = oracle.default_params(5.0, 0.5)              # default friction: static
    m = oracle.Model("box-1x1", "downhill-30", "", oracle.CTRL_PHASE_SIN, p, amp=0.0)
    _, _, tb, _ = m.run(np.zeros(1), trace_steps=300)
    assert np.abs(tb[-1, :, 3:5]).max() < 1e-6 and np.abs(tb[-1, :, :2] - tb[-60, :, :2]).max() < 1e-5


def test_welds_keep_neighbouring_voxels_rigidly_attached(oracle):
    """Robot.assemble() welds the facing corner masses of neighbouring voxels: in the frame of body1 the partner stays at its initial
    offset (one mass side = 0.9 apart, same angle) within the position solver's tolerances (0.005, 2 deg) -- at rest and while the
    voxels are actuated in opposition."""
    for amp in (0.0, 1.0):
        p = oracle.default_params(4.0, 0.5)
        m = oracle.Model("box-2x1", "flat", "", oracle.CTRL_PHASE_SIN, p, amp=amp, freq=1.0)
        _, _, tb, _ = m.run(np.array([0.0, math.pi]), trace_steps=240)
        for a, b in ((1 * 4 + 0, 0 * 4 + 1), (1 * 4 + 3, 0 * 4 + 2)):   # voxel 1's NW / SW on voxel 0's NE / SE
            d = tb[:, a, :3] - tb[:, b, :3]
            assert np.allclose(d[0], [0.9, 0.0, 0.0], atol=1e-12)
            ang = -tb[:, b, 2]
            rx = np.cos(ang) * d[:, 0] - np.sin(ang) * d[:, 1]
            ry = np.sin(ang) * d[:, 0] + np.cos(ang) * d[:, 1]
            assert np.hypot(rx - 0.9, ry).max() < 0.005 and np.abs(d[:, 2]).max() < math.radians(2.0)
