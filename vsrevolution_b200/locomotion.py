"""Host-side mirror of the task boundary: `Locomotion.apply(robot) -> Outcome`, batched.

Mirrors
  Settings                      org.dyn4j.dynamics.Settings                       (locomotion/Starter.java:67)
  Locomotion(finalT, terrain, settings), .apply, createTerrain                    (locomotion/Starter.java:83-94)
  Outcome.getVelocity/getDistance/getTime/subOutcome                              (locomotion/Starter.java:174,430-433)
  buildLocomotionTask                                                             (locomotion/Starter.java:83-94)
plus `apply_all(robots)`: the whole population in one vsr_eval call.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _ffi
from .builder import build_mapper
from .robot import (CentralizedSensing, DistributedSensing, MultiLayerPerceptron, PhaseSin, Robot, RobotUtils, Sinusoid,
                    TimeFunctions)


@dataclass
class Settings:
    """dyn4j Settings() defaults (the only dyn4j symbol the reference touches)."""
    step_frequency: float = 1.0 / 60.0
    velocity_constraint_solver_iterations: int = 10
    position_constraint_solver_iterations: int = 10


@dataclass
class Outcome:
    """What survives of 2dhmsr's Outcome on the device: the observation window [transient_t, final_t)."""
    velocity: float
    distance: float
    time: float
    first_x: float
    last_x: float
    start_t: float
    end_t: float
    control_energy: float = 0.0      # sum over kept steps and voxels of f^2 dt                    [U]
    area_ratio_energy: float = 0.0   # sum over kept steps and voxels of |a_k - a_{k-1}|           [U]
    area_ratio_energy_sq: float = 0.0  # ... of |a_k^2 - a_{k-1}^2| (the other candidate formula)  [U]
    centers: np.ndarray | None = None  # [K][2 or 3] robot centre (x, y[, angle of the central voxel]) at the kept observations (apply_all_observed only)
    dt: float = 1.0 / 60.0
    posture: np.ndarray | None = None     # [n][n] bool, [y][x], y = 0 bottom: Outcome.getAveragePosture(n) (apply_all_observed only)
    footprints: dict | None = None        # {n: uint8 [K]}: bit s = segment s touches the ground (apply_all_observed only)

    # centre spectra of detailedOutcomeFunctions (locomotion/NamedFunctions.java:139-162).  2dhmsr's BehaviorUtils is un-vendored:
    # the definition below is the documented one [U] -- magnitude of the FFT of the series zero-padded to a power of two
    # (commons-math FastFourierTransformer only accepts those), frequencies i / (N dt), averaged inside nBins equal bins of
    # [minF, maxF]; bins without a frequency sample are 0.
    @staticmethod
    def quantized_spectrum(signal: np.ndarray, dt: float, min_f: float, max_f: float, n_bins: int) -> np.ndarray:
        n = 1
        while n < len(signal):
            n *= 2
        mag = np.abs(np.fft.fft(np.concatenate([signal, np.zeros(n - len(signal))])))[: n // 2 + 1]
        f = np.arange(n // 2 + 1) / (n * dt)
        out = np.zeros(n_bins)
        edges = np.linspace(min_f, max_f, n_bins + 1)
        for b in range(n_bins):
            sel = (f >= edges[b]) & ((f < edges[b + 1]) if b + 1 < n_bins else (f <= edges[b + 1]))
            if sel.any():
                out[b] = mag[sel].mean()
        return out

    def _need_centers(self) -> np.ndarray:
        if self.centers is None:
            raise _ffi.VsrError(_ffi.VSR_ESTATE, "centre trajectories were not recorded: use Locomotion.apply_all_observed")
        return self.centers

    def get_center_x_position_spectrum(self, min_f: float, max_f: float, n_bins: int) -> np.ndarray:
        return self.quantized_spectrum(self._need_centers()[:, 0], self.dt, min_f, max_f, n_bins)

    def get_center_x_velocity_spectrum(self, min_f: float, max_f: float, n_bins: int) -> np.ndarray:
        return self.quantized_spectrum(np.diff(self._need_centers()[:, 0]) / self.dt, self.dt, min_f, max_f, n_bins)

    def get_center_y_velocity_spectrum(self, min_f: float, max_f: float, n_bins: int) -> np.ndarray:
        return self.quantized_spectrum(np.diff(self._need_centers()[:, 1]) / self.dt, self.dt, min_f, max_f, n_bins)

    def get_center_angle_spectrum(self, min_f: float, max_f: float, n_bins: int) -> np.ndarray:
        """Outcome.getCenterAngleSpectrum (locomotion/NamedFunctions.java:163-172): spectrum of the central voxel's angle [U]."""
        c = self._need_centers()
        if c.shape[1] < 3:
            raise _ffi.VsrError(_ffi.VSR_ESTATE, "the central voxel's angle was not recorded: use Locomotion.apply_all_observed")
        return self.quantized_spectrum(c[:, 2], self.dt, min_f, max_f, n_bins)

    def get_average_posture(self, n: int) -> np.ndarray:
        """Outcome.getAveragePosture(n) (Starter.java:180-190, NamedFunctions.java:356): [n][n] bool grid, [y][x]."""
        if self.posture is None or self.posture.shape != (n, n):
            raise _ffi.VsrError(_ffi.VSR_ESTATE, f"average posture of side {n} was not recorded: use Locomotion.apply_all_observed(posture_n={n})")
        return self.posture

    def get_footprints_spectra(self, n: int, min_f: float, max_f: float, n_bins: int) -> list[np.ndarray]:
        """Outcome.getFootprintsSpectra(n, minF, maxF, nBins) (locomotion/NamedFunctions.java:179-180,392): one spectrum per footprint
        segment, of the 0/1 series "segment touches the ground" over the kept observations [U]."""
        if self.footprints is None or n not in self.footprints:
            raise _ffi.VsrError(_ffi.VSR_ESTATE, f"footprints of {n} segments were not recorded: use Locomotion.apply_all_observed(footprint_ns=({n},))")
        fp = self.footprints[n]
        return [self.quantized_spectrum(((fp >> s) & 1).astype(np.float64), self.dt, min_f, max_f, n_bins) for s in range(n)]

    # detailedOutcomeFunctions (locomotion/NamedFunctions.java:135-137)
    def get_control_power(self) -> float:
        return self.control_energy / self.time

    def get_area_ratio_power(self) -> float:
        return self.area_ratio_energy / self.time

    def get_corrected_efficiency(self) -> float:
        return self.distance / (1.0 + self.control_energy)

    def get_velocity(self) -> float:
        return self.velocity

    def get_distance(self) -> float:
        return self.distance

    def get_time(self) -> float:
        return self.time

    def sub_outcome(self, start_t: float, end_t: float) -> "Outcome":
        if start_t != self.start_t or end_t != self.end_t:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, "the device only keeps the window given to Locomotion(transient_t=...)")
        return self


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index range of robots owned by `rank`: [rank*ceil(n/world), ...)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class Locomotion:
    """`new Locomotion(finalT, groundProfile, settings)`; `transient_t` folds the reference's
    `.subOutcome(episodeTransientTime, episodeTime)` (Starter.java:430-433) into the same device pass."""

    def __init__(self, final_t: float, ground_profile, settings: Settings | None = None, transient_t: float = 0.0,
                 precision: int = 32, devices: Sequence[int] | None = None):
        self.final_t, self.transient_t = float(final_t), float(transient_t)
        self.xs = np.ascontiguousarray(ground_profile[0], dtype=np.float64)
        self.ys = np.ascontiguousarray(ground_profile[1], dtype=np.float64)
        self.settings = settings or Settings()
        self.precision = precision
        self.devices = list(devices) if devices else []
        self._ctxs: dict = {}

    @staticmethod
    def create_terrain(name: str):
        return _ffi.terrain(name)

    # -- context cache: one vsr_ctx per (body, sensors, controller structure) ---------------------------
    def _context(self, robot: Robot) -> _ffi.Context:
        body = robot.get_voxels()
        mask = body.mask()
        voxels = [v for v in body.values() if v is not None]
        if not voxels:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, "Target robot has no voxels")
        skey = tuple(v.sensor_key() for v in voxels)  # every voxel's own sensor list (they may differ: spinedTouch, px, py, l5)
        c = robot.get_controller()
        if isinstance(c, PhaseSin):
            ckey = ("phase", c.frequency, c.amplitude)
        elif isinstance(c, CentralizedSensing):
            ckey = ("central", tuple(self._mlp(c.function).neurons))
        elif isinstance(c, DistributedSensing):
            fs = [f for f in c.get_functions().values() if f is not None]
            neurons = tuple(self._mlp(fs[0]).neurons)
            if any(tuple(self._mlp(f).neurons) != neurons for f in fs[1:]):
                raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, "per-voxel MLPs of different sizes are outside this build")
            ckey = ("hetero" if self._is_hetero(fs) else "dist", c.signals, neurons)
        elif isinstance(c, TimeFunctions):
            if any(not isinstance(c.get_functions().get(e.key.x, e.key.y), Sinusoid) for e in body if e.value is not None):
                raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, "TimeFunctions: only Sinusoid entries run on the device")
            ckey = ("timesin",)
        else:
            raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, f"controller {type(c).__name__} is outside the hot path")
        key = (mask.tobytes(), mask.shape, skey, ckey)
        ctx = self._ctxs.get(key)
        if ctx is None:
            e = _ffi.env_default()
            e.dt = self.settings.step_frequency
            e.vel_iters = self.settings.velocity_constraint_solver_iterations
            e.pos_iters = self.settings.position_constraint_solver_iterations
            e.final_t, e.transient_t, e.precision = self.final_t, self.transient_t, self.precision
            e.n_devices = len(self.devices)
            for i, d in enumerate(self.devices):
                e.devices[i] = d
            ctx = _ffi.Context(e, self.xs, self.ys)
            off = np.concatenate([[0], np.cumsum([len(v.sensors) for v in voxels])]).astype(np.int32)
            ctx.body_set_voxels(mask, [sd for v in voxels for sd in v.sensors], off)
            if ckey[0] == "phase":
                ctx.controller_set(_ffi.CTRL_PHASE_SIN, (), 0, c.frequency, c.amplitude)
            elif ckey[0] == "central":
                ctx.controller_set(_ffi.CTRL_MLP_CENTRALIZED, ckey[1])
            elif ckey[0] == "timesin":
                ctx.controller_set(_ffi.CTRL_TIME_SIN)
            else:
                kind = _ffi.CTRL_MLP_DISTRIBUTED_HETERO if ckey[0] == "hetero" else _ffi.CTRL_MLP_DISTRIBUTED_HOMO
                ctx.controller_set(kind, ckey[2], c.signals)
            self._ctxs[key] = ctx
        return ctx

    @staticmethod
    def _is_hetero(fs) -> bool:
        """A DistributedSensing whose voxels do not all share one function (BrainHeteroDistributed) needs per-voxel weights."""
        w0 = Locomotion._mlp(fs[0]).get_params()
        return any(f is not fs[0] and not np.array_equal(Locomotion._mlp(f).get_params(), w0) for f in fs[1:])

    @staticmethod
    def _mlp(f) -> MultiLayerPerceptron:
        if not isinstance(f, MultiLayerPerceptron):
            raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, "only MultiLayerPerceptron functions are built (builder/function/MLP.java)")
        return f

    @staticmethod
    def genome_of(robot: Robot) -> np.ndarray:
        """Flatten the controller parameters in the order the kernel expects (== the mappers' genome order)."""
        c = robot.get_controller()
        if isinstance(c, PhaseSin):
            # the phases of the voxels only: brainPhaseFun fills every cell of the grid, voxel or not (BrainPhaseFunction.java:36-43)
            return np.array([c.phases.get(e.key.x, e.key.y) for e in robot.get_voxels() if e.value is not None], dtype=np.float64)
        if isinstance(c, CentralizedSensing):
            return Locomotion._mlp(c.function).get_params()
        if isinstance(c, TimeFunctions):
            body = robot.get_voxels()
            return np.array([[s.frequency, s.phase, s.amplitude] for s in
                             (c.get_functions().get(e.key.x, e.key.y) for e in body if e.value is not None)],
                            dtype=np.float64).reshape(-1)
        fs = [f for f in c.get_functions().values() if f is not None]
        if Locomotion._is_hetero(fs):
            return np.concatenate([Locomotion._mlp(f).get_params() for f in fs])  # grid order == FunctionsGrid's slicing order
        return Locomotion._mlp(fs[0]).get_params()

    def _outcomes(self, res) -> list[Outcome]:
        return [Outcome(float(r["velocity"]), float(r["distance"]), float(r["time"]), float(r["first_x"]), float(r["last_x"]),
                        self.transient_t, self.final_t, float(r["control_energy"]), float(r["area_ratio_abs"]),
                        float(r["area_ratio_sq"])) for r in res]

    def _groups(self, robots: Sequence[Robot]) -> list[tuple[_ffi.Context, list[int]]]:
        """Robots grouped by device context (body, sensors, controller structure), first-appearance order.  A population of one
        morphology is one group = one launch; mappers that also evolve the body (bodyBrainSin) give one launch per distinct
        body (and a homogeneous distributed brain lands in another group than a heterogeneous one)."""
        groups: dict[int, tuple[_ffi.Context, list[int]]] = {}
        for i, r in enumerate(robots):
            ctx = self._context(r)
            groups.setdefault(id(ctx), (ctx, []))[1].append(i)
        return list(groups.values())

    @staticmethod
    def _frozen(robots: Sequence[Robot], idx) -> np.ndarray | None:
        """frozen-actuator masks of the robots a transformation broke (RobotUtils.build_robot_transformation), None if none is"""
        masks = np.array([RobotUtils.frozen_mask(robots[i]) for i in idx], dtype=np.uint32)
        return masks if masks.any() else None

    def _terrains_for(self, n: int):
        return None

    def apply_all(self, robots: Sequence[Robot], terrains=None) -> list[Outcome]:
        """`terrains`: optional per-robot (xs, ys) polylines (what a `-rnd` task draws itself)."""
        if not robots:
            return []
        if terrains is None:
            terrains = self._terrains_for(len(robots))  # -rnd: seeds are drawn in robot order, before any grouping
        out: list = [None] * len(robots)
        for ctx, idx in self._groups(robots):
            genomes = np.stack([self.genome_of(robots[i]) for i in idx])
            frozen = self._frozen(robots, idx)
            if frozen is None and terrains is None:
                _, res = ctx.eval(genomes, want_results=True)
            else:
                _, res, _ = ctx.eval_ex(genomes, frozen_masks=frozen, terrains=None if terrains is None else [terrains[i] for i in idx])
            for i, o in zip(idx, self._outcomes(res)):
                out[i] = o
        return out

    def apply_all_observed(self, robots: Sequence[Robot], posture_n: int = 8, footprint_ns: Sequence[int] = (4, 3)) -> list[Outcome]:
        """apply_all that also keeps, for every robot, what the reference's listeners read from the observations (vsr_eval_ex): the
        centre trajectory + central-voxel angle (centre spectra of `detailedOutcomeFunctions`, NamedFunctions.java:139-172), the
        average posture (`numSpeciated` features Starter.java:180-190, minimap NamedFunctions.java:356) and the footprints
        (NamedFunctions.java:179-180 n = 4, :392 n = 3) -- the whole population in one launch per morphology."""
        if not robots:
            return []
        fns = tuple(footprint_ns) + (0, 0)
        if len(footprint_ns) > 2:
            raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, "at most two footprint segment counts per call")
        terrains = self._terrains_for(len(robots))
        out: list = [None] * len(robots)
        for ctx, idx in self._groups(robots):
            genomes = np.stack([self.genome_of(robots[i]) for i in idx])
            _, res, ex = ctx.eval_ex(genomes, frozen_masks=self._frozen(robots, idx), posture_n=posture_n, footprint_n=fns[:2], centers=True,
                                     terrains=None if terrains is None else [terrains[i] for i in idx])
            for j, (i, o) in enumerate(zip(idx, self._outcomes(res))):
                o.centers, o.dt = ex["centers"][j], self.settings.step_frequency
                if posture_n:
                    o.posture = ex["posture"][j].reshape(posture_n, posture_n).astype(bool)
                o.footprints = {n: ex["footprints"][q][j] for q, n in enumerate(fns[:2]) if n}
                out[i] = o
        return out

    def apply(self, robot: Robot) -> Outcome:
        return self.apply_all([robot])[0]


class JavaRandomInts:
    """The `random.nextInt(bound)` stream of the run's shared java.util.Random (Starter.java:91), replayed by the library."""

    def __init__(self, seed: int):
        self.seed, self.used = seed, 0

    def next_ints(self, bound: int, n: int) -> np.ndarray:
        out = _ffi.java_random_ints(self.seed, bound, self.used + n)[self.used:]
        self.used += n
        return out


class RandomTerrainLocomotion(Locomotion):
    """`r -> new Locomotion(T, createTerrain(name.replace("-rnd", "-" + random.nextInt(10000))), settings).apply(r)`
    (Starter.java:89-93): every evaluation gets its own terrain; the batch goes to the device in one vsr_eval_terrains call.
    Seeds are drawn in robot order (the reference draws them in thread-completion order, which is not reproducible)."""

    def __init__(self, final_t: float, terrain_name: str, random: JavaRandomInts, settings: Settings | None = None,
                 transient_t: float = 0.0, **kw):
        if "-rnd" not in terrain_name:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, "terrain name without -rnd")
        self.terrain_name, self.random = terrain_name, random
        super().__init__(final_t, Locomotion.create_terrain(terrain_name.replace("-rnd", "-0")), settings, transient_t, **kw)

    def terrains_for(self, n: int):
        return [Locomotion.create_terrain(self.terrain_name.replace("-rnd", "-%d" % s)) for s in self.random.next_ints(10000, n)]

    def _terrains_for(self, n: int):
        return self.terrains_for(n)


def build_locomotion_task(terrain_name: str, episode_t: float, transient_t: float = 0.0, random: JavaRandomInts | None = None,
                          **kw) -> Locomotion:
    """Starter.buildLocomotionTask (Starter.java:83-94): one shared Locomotion for fixed terrains, a per-evaluation terrain
    (`random.nextInt(10000)` as seed) for `-rnd` names."""
    if "-rnd" in terrain_name:
        return RandomTerrainLocomotion(episode_t, terrain_name, random or JavaRandomInts(0), Settings(), transient_t, **kw)
    return Locomotion(episode_t, Locomotion.create_terrain(terrain_name), Settings(), transient_t, **kw)


@dataclass
class ValidationOutcome:
    """Starter.ValidationOutcome (locomotion/Starter.java:78-80)."""
    terrain_name: str
    transformation_name: str
    seed: int
    outcome: Outcome


def validate(robots: Sequence[Robot], terrain_name: str, transformation_name: str, seed: int, episode_time: float,
             transient_time: float, **kw) -> list[ValidationOutcome]:
    """Starter.validate (locomotion/Starter.java:122-135) for a batch of robots.  The reference validates ONE robot per call:
    `random = new Random(seed)`, `robot = buildRobotTransformation(name, random).apply(robot)`, then
    `buildLocomotionTask(terrain, T, random, false).apply(robot)` -- so every robot restarts the same stream: the transformation's draws
    (only a `-rnd` seeded transformation takes them from `random`), then the `-rnd` terrain seed `random.nextInt(10000)`.  The batch
    does exactly that per robot and evaluates all of them in one launch per morphology.  Transformations: `identity` (the Starter
    default, :153) and `broken-<ratio>-<seed|rnd>`."""
    transformed, terrains = [], [] if "-rnd" in terrain_name else None
    for r in robots:
        rnd = _ffi.JavaRandomState(seed)
        transformed.append(RobotUtils.build_robot_transformation(transformation_name, rnd)(r))
        if terrains is not None:
            terrains.append(Locomotion.create_terrain(terrain_name.replace("-rnd", "-%d" % rnd.next_int(10000))))
    base = terrain_name.replace("-rnd", "-0")
    task = Locomotion(episode_time, Locomotion.create_terrain(base), Settings(), transient_time, **kw)
    return [ValidationOutcome(terrain_name, transformation_name, seed, o.sub_outcome(transient_time, episode_time))
            for o in task.apply_all(transformed, terrains)]


class MicroBatchedLocomotion:
    """Host strategy (1) of SURVEY.md 8b: keep JGEA's one-robot-per-call contract (`qualityFunction.apply(robot)` from the
    executor's threads, Starter.java:427-436) and batch behind it.  A caller enqueues its robot and blocks; the batch is
    flushed -- by the caller that fills it, or by the first waiter whose `max_wait_s` expires -- with ONE `apply_all` call
    (one kernel launch); every caller then returns its own Outcome.  Needs at least `expected_batch` executor threads to
    fill a batch without waiting for the timeout."""

    class _Batch:
        def __init__(self):
            self.robots, self.outcomes, self.error = [], None, None
            self.done = threading.Event()
            self.t0 = time.monotonic()

    def __init__(self, task: "Locomotion", expected_batch: int, max_wait_s: float = 0.05):
        if expected_batch < 1:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, "expected_batch must be positive")
        self.task, self.expected_batch, self.max_wait_s = task, expected_batch, max_wait_s
        self._lock = threading.Lock()
        self._eval_lock = threading.Lock()  # a vsr_ctx is not re-entrant
        self._open = self._Batch()
        self.batches_flushed = 0

    def _flush(self, batch: "_Batch"):
        try:
            with self._eval_lock:
                batch.outcomes = self.task.apply_all(batch.robots)
        except Exception as e:  # every waiter sees the failure, like an exception thrown by Locomotion.apply
            batch.error = e
        self.batches_flushed += 1
        batch.done.set()

    def apply(self, robot: Robot) -> Outcome:
        with self._lock:
            batch = self._open
            slot = len(batch.robots)
            batch.robots.append(robot)
            full = len(batch.robots) >= self.expected_batch
            if full:
                self._open = self._Batch()
        if full:
            self._flush(batch)
        else:
            while not batch.done.wait(max(0.0, batch.t0 + self.max_wait_s - time.monotonic())):
                with self._lock:  # timeout: the first waiter that still finds the batch open closes and flushes it
                    leader = self._open is batch
                    if leader:
                        self._open = self._Batch()
                if leader:
                    self._flush(batch)
                else:
                    batch.done.wait()
        if batch.error is not None:
            raise batch.error
        return batch.outcomes[slot]


class LazyOutcome:
    """What `LazyBatchedLocomotion.apply` returns: an Outcome whose getters are answered after the whole batch it belongs to has been
    evaluated in ONE launch; the first getter called on any outcome of the batch triggers that evaluation
    (java/it/units/erallab/gpu/GpuOutcome.java is the Java form)."""

    def __init__(self, batch: "LazyBatchedLocomotion._Batch", slot: int):
        self._batch, self._slot = batch, slot

    def _resolved(self) -> Outcome:
        return self._batch.resolve()[self._slot]

    def sub_outcome(self, start_t: float, end_t: float) -> "LazyOutcome":
        task = self._batch.task
        if start_t != task.transient_t or end_t != task.final_t:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, "the device only keeps the window given to Locomotion(transient_t=...)")
        return self  # still lazy: Starter.java:430-433 calls it on every outcome before anybody reads a value

    def __getattr__(self, name):  # every other getter / field of Outcome: resolve, then delegate
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._resolved(), name)


class LazyBatchedLocomotion:
    """Host strategy (3) of SURVEY.md 8b, the one the Java shim uses by default: keep JGEA's one-robot-per-call contract
    (`qualityFunction.apply(robot)`, Starter.java:427-436) with NO extra threads and NO time-outs.  `apply(robot)` only records the
    robot and returns a LazyOutcome; JGEA builds its Individuals without reading the fitness object and reads it when it orders the new
    population -- at that moment the whole generation goes to the device in one `apply_all` call."""

    class _Batch:
        def __init__(self, task):
            self.task, self.robots, self.outcomes, self.error = task, [], None, None
            self.lock = threading.Lock()

        def resolve(self):
            with self.lock:
                if self.outcomes is None and self.error is None:
                    self.task._close(self)
                    try:
                        self.outcomes = self.task._evaluate(self.robots)
                    except Exception as e:  # every outcome of the batch reports the failure, like Locomotion.apply throwing
                        self.error = e
            if self.error is not None:
                raise self.error
            return self.outcomes

    def __init__(self, task: "Locomotion", observed: bool = False):
        self.task, self.observed = task, observed
        self.transient_t, self.final_t = task.transient_t, task.final_t
        self._lock = threading.Lock()
        self._open = self._Batch(self)
        self.batches_flushed = 0

    def apply(self, robot: Robot) -> LazyOutcome:
        with self._lock:
            batch = self._open
            batch.robots.append(robot)
            return LazyOutcome(batch, len(batch.robots) - 1)

    def _close(self, batch):
        with self._lock:
            if self._open is batch:
                self._open = self._Batch(self)  # robots applied from now on belong to the next batch

    def _evaluate(self, robots):
        self.batches_flushed += 1
        return self.task.apply_all_observed(robots) if self.observed else self.task.apply_all(robots)


class PopulationEvaluator:
    """genome matrix -> fitness vector for one (shape, sensors, mapper, terrain): what a batching JGEA evolver calls
    once per generation.  For the mappers whose genome IS the controller parameter vector (brainCentralized<mlp,
    brainPhaseVals, brainHomoDist<mlp, brainHeteroDist<funsGrid) rows go to the device untouched (`direct`); the mappers that
    transform the genome on the host (brainPhaseFun<mlp, bodyBrainSin<dirNumGrid ...) are applied row by row and the
    robots go through `apply_all`, which groups them by morphology."""

    def __init__(self, shape: str, sensor_config: str, mapper: str, terrain: str, episode_t: float, transient_t: float,
                 precision: int = 32, devices: Sequence[int] | None = None, transformation: str = "identity", seed: int = 0):
        self.task = build_locomotion_task(terrain, episode_t, transient_t, precision=precision, devices=devices)
        # `transformation.andThen(task)` of Problem.qualityFunction (Starter.java:427-429): one Random for the whole run
        self.transformation_name = transformation
        self._t_random = _ffi.JavaRandomState(seed)
        self._t_spec = _ffi.robot_transformation(transformation)
        self._transform = RobotUtils.build_robot_transformation(transformation, self._t_random)
        body = RobotUtils.build_sensorizing_function(sensor_config)(RobotUtils.build_shape(shape))
        self.target = Robot(None, body)
        self.mapper = build_mapper(mapper)
        self.genome_len = len(self.mapper.example_for(self.target))
        self._build = self.mapper.build_for(self.target)
        # direct <=> a probe genome comes back unchanged as the device genome of the robot it builds (distinct values, so a
        # heterogeneous brain is recognised as such)
        probe = np.linspace(-0.5, 0.5, self.genome_len)
        proto = self._build(probe)
        self.ctx = self.task._context(proto)
        dev = Locomotion.genome_of(proto)
        self.direct = self.ctx.genome_len == self.genome_len and np.array_equal(dev, probe)
        self.n_voxels = self.ctx.n_voxels

    def evaluate(self, genomes: np.ndarray) -> np.ndarray:
        if not self.direct:
            return np.array([o.get_velocity() for o in
                             self.task.apply_all([self._transform(self._build(g)) for g in np.asarray(genomes)])])
        n = len(genomes)
        frozen = None
        if self._t_spec[0] == "broken":  # the library replays the transformation's draws on the run's stream, in robot order
            frozen = self.ctx.broken_masks_stream(self._t_spec[1], self._transform_state(), n)
        terrains = self.task.terrains_for(n) if isinstance(self.task, RandomTerrainLocomotion) else None
        if frozen is None and terrains is None:
            return self.ctx.eval(genomes)
        return self.ctx.eval_ex(genomes, frozen_masks=frozen, terrains=terrains)[0]

    @property
    def has_per_evaluation_inputs(self) -> bool:
        return self.direct and (self._t_spec[0] == "broken" or isinstance(self.task, RandomTerrainLocomotion))

    def per_evaluation_inputs(self, n: int):
        """(frozen masks or None, terrains or None) for the next n evaluations, drawn in evaluation order: what the device-resident
        generation loop stages per batch (solver.GenerationLoop)."""
        frozen = self.ctx.broken_masks_stream(self._t_spec[1], self._transform_state(), n) if self._t_spec[0] == "broken" else None
        terrains = self.task.terrains_for(n) if isinstance(self.task, RandomTerrainLocomotion) else None
        return frozen, terrains

    def _transform_state(self) -> "_ffi.JavaRandomState":
        """the Random the `broken` transformation draws from: its own `new Random(seed)` (created once, at build time), or the run's"""
        if self._t_spec[2] is not None:
            if not hasattr(self, "_own_random"):
                self._own_random = _ffi.JavaRandomState(self._t_spec[2])
            return self._own_random
        return self._t_random

    def kernel_launches(self) -> int:
        return self.ctx.kernel_launches()


def population_evaluator(shape: str, sensor_config: str, mapper: str, terrain: str, episode_t: float,
                         transient_t: float, **kw) -> PopulationEvaluator:
    return PopulationEvaluator(shape, sensor_config, mapper, terrain, episode_t, transient_t, **kw)
