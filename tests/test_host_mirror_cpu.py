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
        B.build_mapper("brainHeteroDist<mlp")


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
