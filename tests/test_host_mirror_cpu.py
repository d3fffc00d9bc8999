"""Host-side mirror of the reference mappers: same dimension checks, same genome layout (no GPU needed)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def V():
    import __graft_entry__ as g
    import os
    from vsrevolution_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        g.build()
    import vsrevolution_b200.builder as B
    import vsrevolution_b200.locomotion as L
    import vsrevolution_b200.robot as R
    return B, L, R, _ffi


def _target(R, shape, sensors):
    return R.Robot(None, R.RobotUtils.build_sensorizing_function(sensors)(R.RobotUtils.build_shape(shape)))


def test_named_provider_parse(V):
    B, *_ = V
    assert B.NamedProvider.parse("mlp;r=0.65;nIL=1") == {"NAME": "mlp", "r": "0.65", "nIL": "1"}
    assert B.MAPPER_PROVIDER.build("nope;x=1") is None


def test_brain_centralized_mlp_genome_length_and_checks(V):
    B, L, R, F = V
    t = _target(R, "biped-4x3", "uniform-a-0")
    mapper = B.build_mapper("brainCentralized<mlp;r=0.65;nIL=1")
    assert len(mapper.example_for(t)) == 157
    robot = mapper.build_for(t)(np.arange(157) / 157.0)
    assert isinstance(robot.get_controller(), R.CentralizedSensing)
    assert robot.get_controller().function.neurons == [10, 7, 10]
    assert np.array_equal(L.Locomotion.genome_of(robot), np.arange(157) / 157.0)
    with pytest.raises(ValueError, match="Wrong number of values for weights: 157 expected, 156 found"):
        mapper.build_for(t)(np.zeros(156))
    # wrong function shape -> the reference's message (BrainCentralized.java:28-33)
    with pytest.raises(ValueError, match="Wrong number of function input args: 10 expected, 9 found"):
        B.BrainCentralized().build({}).build_for(t)(R.RealFunction(9, 10))


def test_brain_phase_values_order(V):
    B, L, R, F = V
    t = _target(R, "biped-4x3", "")
    mapper = B.build_mapper("brainPhaseVals;f=0.5;a=0.7")
    assert len(mapper.example_for(t)) == 10
    robot = mapper.build_for(t)(list(range(10)))
    c = robot.get_controller()
    assert (c.frequency, c.amplitude) == (0.5, 0.7)
    # c-th value -> c-th non-null voxel in Grid iteration order (row-major from the bottom row)
    assert c.phases.get(0, 0) == 0 and c.phases.get(3, 0) == 1 and c.phases.get(1, 0) is None and c.phases.get(0, 1) == 2
    assert np.array_equal(L.Locomotion.genome_of(robot), np.arange(10.0))
    with pytest.raises(ValueError, match="Wrong number of values: 10 expected, 3 found"):
        mapper.build_for(t)([0, 1, 2])


def test_brain_homo_distributed(V):
    B, L, R, F = V
    t = _target(R, "comb-7x2", "uniform-t+vxy+a-0")
    mapper = B.build_mapper("brainHomoDist;s=1<mlp;r=0.65;nIL=1")
    assert len(mapper.example_for(t)) == 75
    robot = mapper.build_for(t)(np.linspace(-1, 1, 75))
    c = robot.get_controller()
    assert c.signals == 1 and sum(f is not None for f in c.get_functions().values()) == 11
    assert np.array_equal(L.Locomotion.genome_of(robot), np.linspace(-1, 1, 75))


def test_shard_range_partitions(V):
    _, L, *_ = V
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            parts = [L.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def test_unsupported_is_loud(V):
    B, L, R, F = V
    with pytest.raises(F.VsrError):
        B.build_mapper("brainAutoPoses;nP=4")  # PosesController: outside the hot path
    with pytest.raises(F.VsrError):
        B.build_mapper("sensorBrainCentralized")


def test_brain_hetero_distributed_functions_grid(V):
    """brainHeteroDist<funsGrid (BrainHeteroDistributed.java:36-83, FunctionsGrid.java:31-46): the flat genome is sliced into one
    MLP per non-null voxel in Grid iteration order; the device genome is the same vector."""
    B, L, R, F = V
    t = _target(R, "comb-7x2", "uniform-t+vxy+a-0")
    mapper = B.build_mapper("brainHeteroDist;s=1<funsGrid")
    assert len(mapper.example_for(t)) == 11 * 75
    g = np.linspace(-1, 1, 825)
    robot = mapper.build_for(t)(g)
    c = robot.get_controller()
    fs = c.get_functions()
    assert isinstance(c, R.DistributedSensing) and c.signals == 1
    # comb-7x2: full top row (y = 1), even-x bottom row (y = 0); iteration = bottom row first
    assert fs.get(1, 0) is None and np.array_equal(fs.get(0, 0).get_params(), g[:75]) and np.array_equal(fs.get(2, 0).get_params(), g[75:150])
    assert np.array_equal(fs.get(0, 1).get_params(), g[4 * 75:5 * 75])
    assert np.array_equal(L.Locomotion.genome_of(robot), g)
    # a functions grid of the wrong size / arity gives the reference's messages
    bf = B.BrainHeteroDistributed().build({"s": "1"}).build_for(t)
    with pytest.raises(ValueError, match="Wrong size of functions grid: 7x2 expected, 3x2 found"):
        bf(R.Grid(3, 2, R.RealFunction(8, 5)))
    with pytest.raises(ValueError, match=r"Wrong number of function input args at \(0,0\): 8 expected, 7 found"):
        bf(R.Grid(7, 2, R.RealFunction(7, 5)))
    with pytest.raises(ValueError, match=r"Wrong number of function output args at \(0,0\): 5 expected, 4 found"):
        bf(R.Grid(7, 2, R.RealFunction(8, 4)))


def test_brain_phase_function_divides_both_coordinates_by_the_width(V, oracle):
    """brainPhaseFun<mlp (BrainPhaseFunction.java:36-43): phase(x, y) = f(0, [x / w, y / w])[0] for every cell."""
    B, L, R, F = V
    t = _target(R, "biped-4x3", "")
    mapper = B.build_mapper("brainPhaseFun;f=0.5;a=0.9<mlp;r=0.65;nIL=1")
    G = len(mapper.example_for(t))
    assert G == 9  # 2-2-1: innerNeurons(2, 1, 0.65, 1) = [max(round(2 * 0.65), 2)]... evaluated by the C library
    w = np.random.default_rng(4).uniform(-1, 1, G)
    robot = mapper.build_for(t)(w)
    c = robot.get_controller()
    assert (c.frequency, c.amplitude) == (0.5, 0.9)
    mlp = R.MultiLayerPerceptron("tanh", 2, F.mlp_inner_neurons(2, 1, 0.65, 1), 1, w)
    for (x, y) in ((0, 0), (3, 0), (1, 0), (2, 2)):
        assert c.phases.get(x, y) == mlp.apply([x / 4, y / 4])[0]   # y / 4, not y / 3; (1, 0) is not a voxel but still has a phase
    dev = L.Locomotion.genome_of(robot)
    assert len(dev) == 10 and dev[1] == c.phases.get(3, 0) and dev[2] == c.phases.get(0, 1)
    with pytest.raises(ValueError, match="Wrong number of function input args: 2 expected, 3 found"):
        B.BrainPhaseFunction().build({}).build_for(t)(R.RealFunction(3, 1))


def test_host_mlp_apply_is_the_oracle_mlp(V, oracle):
    """MultiLayerPerceptron.apply on the host == the oracle's controller arithmetic: a 1x1 robot whose centralized MLP maps its
    area-ratio reading to the control value; first-step control from the oracle trace vs the host evaluation."""
    B, L, R, F = V
    m = oracle.Model("box-1x1", "flat", "uniform-a-0", oracle.CTRL_MLP_CENTRALIZED, oracle.default_params(0.5, 0.0), n_inner=1)
    G = m.genome_len()
    g = np.random.default_rng(8).uniform(-1, 1, G)
    _, _, _, tc = m.run(g, trace_steps=1)
    layers = [m.ctrl.layers[i] for i in range(m.ctrl.n_layers)]
    mlp = R.MultiLayerPerceptron("tanh", layers[0], layers[1:-1], layers[-1], g)
    # the reading at step 1: soft-normalised area ratio of the (still almost undeformed) voxel; recover it by inverting nothing --
    # evaluate the MLP over a sweep and check the oracle's control value is attained at a reading close to the rest value
    xs = np.linspace(0.45, 0.55, 20001)
    ys = np.array([mlp.apply([x])[0] for x in xs[::200]])
    assert ys.min() - 1e-6 <= tc[0, 0] <= ys.max() + 1e-6


def test_direct_numbers_grid_is_column_major(V):
    B, L, R, F = V
    grid = R.Grid(3, 2, lambda x, y: [0.0, 0.0])
    out = B.DirectNumbersGrid().build({}).build_for(grid)(list(range(12)))
    assert out.get(0, 0) == [0, 1] and out.get(0, 1) == [2, 3] and out.get(1, 0) == [4, 5] and out.get(2, 1) == [10, 11]
    with pytest.raises(ValueError, match="Wrong number of values: 12 expected, 11 found"):
        B.DirectNumbersGrid().build({}).build_for(grid)(list(range(11)))


def test_percentile_is_commons_math_legacy(V):
    B, *_ = V
    # documented examples of the LEGACY estimator: pos = p (n + 1) / 100
    assert B.percentile_legacy([1, 2, 3, 4, 5], 50) == 3.0
    assert B.percentile_legacy([1, 2, 3, 4], 50) == 2.5
    assert B.percentile_legacy([4, 1, 3, 2], 25) == 1.25
    assert B.percentile_legacy([1, 2, 3, 4], 99) == 4.0
    assert B.percentile_legacy([7.0], 30) == 7.0
    # the Starter's default p = 0.5 is half a PERCENT: on any grid below 199 cells the threshold is the minimum
    assert B.percentile_legacy(list(range(12)), 0.5) == 0.0


def test_body_brain_sinusoidal(V):
    """bodyBrainSin<dirNumGrid (BodyBrainSinusoidal.java:36-124): with the default p the whole grid becomes the body; with p = 50
    the upper half by body value, largest connected component, cropped; frequencies / phases decoded from clipped genes."""
    B, L, R, F = V
    t = _target(R, "biped-4x3", "")
    mapper = B.build_mapper("bodyBrainSin<dirNumGrid")
    G = len(mapper.example_for(t))
    assert G == 4 * 3 * 3
    vals = np.zeros((4, 3, 3))  # [x][y][(body, frequency, phase)] in dirNumGrid's column-major order
    vals[..., 0] = np.arange(12).reshape(4, 3)
    vals[..., 1] = -1.0
    vals[..., 2] = 3.0  # clipped to 1
    vals[1, 2, 1] = 0.0
    robot = mapper.build_for(t)(vals.reshape(-1))
    assert robot.voxels.mask().all() and robot.voxels.w == 4 and robot.voxels.h == 3
    fns = robot.get_controller().get_functions()
    assert fns.get(0, 0) == R.Sinusoid(0.5, np.pi, 1.0) and fns.get(1, 2) == R.Sinusoid(1.25, np.pi, 1.0)
    dev = L.Locomotion.genome_of(robot).reshape(12, 3)
    assert np.array_equal(dev[1 + 2 * 4], [1.25, np.pi, 1.0])  # grid order: y * w + x
    half = B.build_mapper("bodyBrainSin;p=50<dirNumGrid").build_for(t)(vals.reshape(-1))
    # body values 0..11 column-major: threshold = 5.5 -> x = 2, 3 columns (values 6..11) survive, cropped to 2x3
    assert half.voxels.w == 2 and half.voxels.h == 3 and half.voxels.mask().all()
    with pytest.raises(ValueError, match="Wrong grid size: 4x3 expected, 3x3 found"):
        B.BodyBrainSinusoidal().build({}).build_for(t)(R.Grid(3, 3, lambda x, y: [0.0] * 3))
    with pytest.raises(ValueError, match="Wrong number of values in grid at 0,0: 3 expected, 2 found"):
        B.BodyBrainSinusoidal().build({}).build_for(t)(R.Grid(4, 3, lambda x, y: [0.0] * 2))


def test_apply_all_groups_robots_by_morphology(V):
    """Mixed populations (body-evolving mappers): one device batch per distinct (body, sensors, controller structure), results
    returned in the callers' order.  Runs without a GPU against stand-in contexts."""
    B, L, R, F = V
    t = _target(R, "biped-4x3", "")
    full = B.build_mapper("bodyBrainSin<dirNumGrid").build_for(t)
    half = B.build_mapper("bodyBrainSin;p=50<dirNumGrid").build_for(t)
    rng = np.random.default_rng(0)
    robots = [(full if i % 3 else half)(rng.uniform(-1, 1, 36)) for i in range(7)]

    class FakeCtx:
        def __init__(self, nv):
            self.nv, self.calls = nv, []
        def eval(self, genomes, want_results=False):
            self.calls.append(genomes.shape)
            res = np.zeros(len(genomes), dtype=F.RESULT_DTYPE)
            res["velocity"] = genomes[:, 0] + 100 * self.nv
            res["time"] = 1.0
            return res["velocity"].copy(), res

    task = L.Locomotion(5.0, ([0.0, 10.0], [0.0, 0.0]))
    fakes = {}
    task._context = lambda r: fakes.setdefault(int(r.voxels.mask().sum()), FakeCtx(int(r.voxels.mask().sum())))
    out = task.apply_all(robots)
    assert sorted(fakes) == sorted({int(r.voxels.mask().sum()) for r in robots}) and len(fakes) >= 2
    assert all(len(c.calls) == 1 for c in fakes.values())
    for r, o in zip(robots, out):
        nv = int(r.voxels.mask().sum())
        assert o.get_velocity() == L.Locomotion.genome_of(r)[0] + 100 * nv


def test_rnd_terrain_task_draws_java_seeds(V, oracle):
    """Starter.java:89-93: `name.replace("-rnd", "-" + random.nextInt(10000))` per evaluation, from the run's java.util.Random."""
    B, L, R, F = V
    task = L.build_locomotion_task("hilly-1-10-rnd", 10.0, 1.0, random=L.JavaRandomInts(5))
    assert isinstance(task, L.RandomTerrainLocomotion)
    r = oracle.JavaRandom(5)
    seeds = [r.next_int(10000) for _ in range(5)]
    got = task.terrains_for(3) + task.terrains_for(2)  # the stream continues across batches
    for s, (xs, ys) in zip(seeds, got):
        oxs, oys = oracle.create_terrain("hilly-1-10-%d" % s)
        assert np.array_equal(xs, oxs) and np.array_equal(ys, oys)
    assert F.java_random_ints(5, 10000, 5).tolist() == seeds
    rp = oracle.JavaRandom(-3)
    assert F.java_random_ints(-3, 16, 4).tolist() == [rp.next_int(16) for _ in range(4)]  # power-of-two bound path
    with pytest.raises(F.VsrIllegalArgument):
        L.RandomTerrainLocomotion(10.0, "flat", L.JavaRandomInts(0))


def test_micro_batching_keeps_the_one_robot_per_call_contract(V):
    """SURVEY 8b strategy (1): many threads call apply(robot); one apply_all per batch; timeout flush for a partial batch;
    every caller gets its own result; an evaluator error reaches every waiter."""
    import threading
    B, L, R, F = V

    class FakeTask:  # stands in for Locomotion on a machine without a GPU
        def __init__(self):
            self.calls = []
        def apply_all(self, robots):
            self.calls.append(len(robots))
            if any(r == "boom" for r in robots):
                raise RuntimeError("device failure")
            return [("outcome", r) for r in robots]

    task = FakeTask()
    mb = L.MicroBatchedLocomotion(task, expected_batch=16, max_wait_s=0.2)
    out = {}
    def call(i):
        out[i] = mb.apply(i)
    ts = [threading.Thread(target=call, args=(i,)) for i in range(32)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert sorted(task.calls) == [16, 16] and all(out[i] == ("outcome", i) for i in range(32))
    # partial batch: flushed by the timeout, once
    task.calls.clear(); out.clear()
    ts = [threading.Thread(target=call, args=(i,)) for i in range(5)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert task.calls == [5] and all(out[i] == ("outcome", i) for i in range(5))
    # errors propagate to every caller of the batch
    errs = []
    def call_bad(x):
        try:
            mb.apply(x)
        except RuntimeError as e:
            errs.append(str(e))
    ts = [threading.Thread(target=call_bad, args=(x,)) for x in ("a", "boom", "c")]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert errs == ["device failure"] * 3
    with pytest.raises(F.VsrIllegalArgument):
        L.MicroBatchedLocomotion(task, 0)


def test_outcome_quantized_spectrum(V):
    """The centre spectra of detailedOutcomeFunctions (NamedFunctions.java:139-162) on a synthetic trajectory: a 2.25 Hz oscillation
    of the centre velocity lands in the bin that holds it; without recorded centres the getters fail loudly."""
    B, L, R, F = V
    dt = 1 / 60.0
    t = np.arange(1, 541) * dt
    x = 0.5 * t + 0.02 * np.sin(2 * np.pi * 2.25 * t)
    y = 1.0 + 0.01 * np.sin(2 * np.pi * 3.25 * t)
    o = L.Outcome(0.5, 4.5, 9.0, 0.0, 4.5, 1.0, 10.0, centers=np.stack([x, y], axis=1), dt=dt)
    sx = o.get_center_x_velocity_spectrum(0.0, 4.0, 8)   # Starter.java:139-141: [0, 4] Hz, spectrumSize bins
    sy = o.get_center_y_velocity_spectrum(0.0, 4.0, 8)
    assert len(sx) == 8 and int(np.argmax(sx[1:])) + 1 == 4 and int(np.argmax(sy)) == 6   # 2.25 Hz -> bin [2, 2.5), 3.25 Hz -> bin [3, 3.5)
    assert int(np.argmax(o.get_center_x_position_spectrum(0.0, 4.0, 8))) == 0             # the drift dominates the position spectrum
    with pytest.raises(F.VsrError):
        L.Outcome(0.5, 4.5, 9.0, 0.0, 4.5, 1.0, 10.0).get_center_x_velocity_spectrum(0, 4, 8)


def test_lazy_batching_evaluates_a_generation_in_one_call(V):
    """LazyBatchedLocomotion (the Java shim's default strategy): apply() records, the first getter on any outcome of the batch evaluates
    the whole batch once; robots applied afterwards open the next batch; an evaluation error reaches every outcome of its batch."""
    from concurrent.futures import ThreadPoolExecutor
    from vsrevolution_b200.locomotion import LazyBatchedLocomotion, Outcome

    class FakeTask:
        transient_t, final_t = 1.0, 3.0

        def __init__(self):
            self.calls = []

        def apply_all(self, robots):
            self.calls.append(len(robots))
            if any(r == "boom" for r in robots):
                raise RuntimeError("device failure")
            return [Outcome(float(r), float(r) * 2, 2.0, 0.0, 0.0, 1.0, 3.0) for r in robots]

    task = FakeTask()
    lazy = LazyBatchedLocomotion(task)
    with ThreadPoolExecutor(8) as ex:                                 # JGEA's executor threads (Starter.java:436)
        gen1 = list(ex.map(lambda r: lazy.apply(r).sub_outcome(1.0, 3.0), range(40)))
    assert task.calls == []                                            # nothing evaluated while the Individuals are being built
    ranked = sorted(gen1, key=lambda o: -o.get_velocity())             # the comparator reads the first fitness: one call for all 40
    assert task.calls == [40] and ranked[0].get_velocity() == 39.0 and lazy.batches_flushed == 1
    assert sorted(o.get_distance() for o in gen1) == [2.0 * i for i in range(40)]
    gen2 = [lazy.apply(r) for r in (5, 6, 7)]
    assert gen2[2].get_velocity() == 7.0 and task.calls == [40, 3]
    bad = [lazy.apply(r) for r in (1, "boom", 2)]
    for o in bad:
        with pytest.raises(RuntimeError):
            o.get_velocity()
    assert task.calls == [40, 3, 3]
    with pytest.raises(ValueError):
        lazy.apply(1).sub_outcome(0.0, 3.0)
