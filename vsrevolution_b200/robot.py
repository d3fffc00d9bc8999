"""Host-side mirror of the 2dhmsr types the reference hands to the hot path (Robot = body grid + sensors +
controller).  These are plain descriptions: all simulation happens in libvsr_eval.so.

Mirrors (names and argument meaning) of, as used by the reference:
  Grid                      it.units.erallab.hmsrobots.util.Grid            (builder/robot/BrainPhaseValues.java:38-46)
  Voxel(sensors)            ...core.objects.Voxel                           (builder/robot/SensorBrainCentralized.java:81-87)
  Robot(controller, voxels) ...core.objects.Robot                           (builder/robot/BrainCentralized.java:42-45)
  PhaseSin(f, a, phases)    ...core.controllers.PhaseSin                    (builder/robot/BrainPhaseValues.java:47-50)
  CentralizedSensing        ...core.controllers.CentralizedSensing          (builder/robot/BrainCentralized.java:28-43)
  DistributedSensing        ...core.controllers.DistributedSensing          (builder/robot/BrainHomoDistributed.java:76-82,
                                                                             builder/robot/BrainHeteroDistributed.java:76-83)
  TimeFunctions             ...core.controllers.TimeFunctions               (builder/robot/BodyBrainSinusoidal.java:87-119)
  MultiLayerPerceptron      ...core.controllers.MultiLayerPerceptron        (builder/function/MLP.java:63-69)
  RealFunction.build        ...core.controllers.RealFunction                (builder/robot/BrainCentralized.java:52-56)
  RobotUtils.buildShape / buildSensorizingFunction                          (locomotion/Starter.java:402-406)
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Callable, Generic, Iterator, Optional, Sequence, TypeVar

import numpy as np

from . import _ffi

T = TypeVar("T")


class Grid(Generic[T]):
    """w x h grid; get(x, y) = ts[y*w + x]; values() and iteration are row-major (y outer, x inner), y=0 bottom."""

    @dataclass(frozen=True)
    class Key:
        x: int
        y: int

    @dataclass(frozen=True)
    class Entry(Generic[T]):
        key: "Grid.Key"
        value: T

    def __init__(self, w: int, h: int, filler=None):
        self.w, self.h = w, h
        if callable(filler):
            self._ts = [filler(i % w, i // w) for i in range(w * h)]
        else:
            self._ts = [filler for _ in range(w * h)]

    @staticmethod
    def create(other: "Grid", fn: Optional[Callable] = None) -> "Grid":
        g = Grid(other.w, other.h)
        if fn is not None:
            g._ts = [fn(v) for v in other._ts]
        return g

    def get(self, x: int, y: int):
        if x < 0 or x >= self.w or y < 0 or y >= self.h:
            return None
        return self._ts[y * self.w + x]

    def set(self, x: int, y: int, v) -> None:
        self._ts[y * self.w + x] = v

    def values(self) -> list:
        return list(self._ts)

    def __iter__(self) -> Iterator["Grid.Entry"]:
        for i, v in enumerate(self._ts):
            yield Grid.Entry(Grid.Key(i % self.w, i // self.w), v)

    def count(self, pred: Callable) -> int:
        return sum(1 for v in self._ts if pred(v))

    def mask(self) -> np.ndarray:
        return np.array([0 if v is None or v is False else 1 for v in self._ts], dtype=np.uint8).reshape(self.h, self.w)


@dataclass
class Voxel:
    sensors: list = field(default_factory=list)  # list[_ffi.SensorDesc]

    def sensor_dim(self) -> int:
        return sum(s.n_rays if s.kind == _ffi.SENSOR_LIDAR else (bin(s.axes & 3).count("1") if s.kind == _ffi.SENSOR_VELOCITY else 1)
                   for s in self.sensors)

    def sensor_key(self) -> tuple:
        return tuple(s.key() for s in self.sensors)


@dataclass
class BreakableVoxel(Voxel):
    """2dhmsr BreakableVoxel as `broken-<ratio>-<seed>` builds it [U]: ACTUATOR -> {FROZEN}, trigger TIME >= 0, restore time
    infinite, i.e. the actuator applies its last force before freezing (none) for the whole episode; sensors keep working."""
    seed: int = 0
    actuator_frozen: bool = True


# ---- functions ---------------------------------------------------------------------------------------
class TimedRealFunction:
    def get_input_dimension(self) -> int:
        raise NotImplementedError

    def get_output_dimension(self) -> int:
        raise NotImplementedError


class RealFunction(TimedRealFunction):
    """Shape-only placeholder, RealFunction.build(f, nIn, nOut) in the reference's exampleFor()."""

    def __init__(self, n_in: int, n_out: int):
        self.n_in, self.n_out = n_in, n_out

    @staticmethod
    def build(_f, n_in: int, n_out: int) -> "RealFunction":
        return RealFunction(n_in, n_out)

    def get_input_dimension(self) -> int:
        return self.n_in

    def get_output_dimension(self) -> int:
        return self.n_out


class MultiLayerPerceptron(TimedRealFunction):
    TANH = "tanh"

    def __init__(self, activation: str, n_in: int, inner: Sequence[int], n_out: int, weights: Sequence[float]):
        if activation != self.TANH:
            raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, "only the TANH activation (builder/function/MLP.java:20-22) is built")
        self.neurons = [n_in, *inner, n_out]
        self.weights = np.asarray(weights, dtype=np.float64)
        n = self.count_weights(self.neurons)
        if n != len(self.weights):
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EDIM, f"Wrong number of values for weights: {n} expected, {len(self.weights)} found")

    @staticmethod
    def count_neurons(n_in: int, inner: Sequence[int], n_out: int) -> list[int]:
        return [n_in, *inner, n_out]

    @staticmethod
    def count_weights(neurons_or_in, inner=None, n_out=None) -> int:
        neurons = neurons_or_in if inner is None else [neurons_or_in, *inner, n_out]
        return _ffi.mlp_count_weights(list(neurons))

    def get_input_dimension(self) -> int:
        return self.neurons[0]

    def get_output_dimension(self) -> int:
        return self.neurons[-1]

    def get_params(self) -> np.ndarray:
        return self.weights

    def apply(self, t_or_input, input_=None, input_activation: bool = True) -> np.ndarray:
        """`function.apply(t, in)` / `function.apply(in)` on the host, in double: used by the mappers that evaluate a function
        once per voxel at build time (BrainPhaseFunction.java:37-43, FunctionNumbersGrid.java:50-54), never per step.  Same
        arithmetic as the oracle's mlp_apply (weights layer -> to-neuron -> (bias, from-neurons...), tanh on every layer)."""
        x = np.asarray(t_or_input if input_ is None else input_, dtype=np.float64)
        if len(x) != self.neurons[0]:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EDIM, f"Expected input length is {self.neurons[0]}: found {len(x)}")
        if input_activation:
            x = np.tanh(x)
        o = 0
        for l in range(1, len(self.neurons)):
            nf, nt = self.neurons[l - 1], self.neurons[l]
            w = self.weights[o:o + nt * (nf + 1)].reshape(nt, nf + 1)
            y = np.empty(nt)
            for j in range(nt):  # same summation order as the oracle: bias, then the from-neurons in order
                acc = w[j, 0]
                for k in range(nf):
                    acc += x[k] * w[j, 1 + k]
                y[j] = np.tanh(acc)
            x = y
            o += nt * (nf + 1)
        return x


# ---- controllers ---------------------------------------------------------------------------------------
class Controller:
    @staticmethod
    def empty() -> "Controller":
        return Controller()


class PhaseSin(Controller):
    def __init__(self, frequency: float, amplitude: float, phases: Grid):
        self.frequency, self.amplitude, self.phases = frequency, amplitude, phases


@dataclass(frozen=True)
class Sinusoid:
    """`t -> amplitude * Math.sin(2 * Math.PI * freq * t + phase)` (BodyBrainSinusoidal.java:115): the only time function the
    reference builds for TimeFunctions; kept as data so that the device can evaluate it."""
    frequency: float
    phase: float
    amplitude: float

    def __call__(self, t: float) -> float:
        return self.amplitude * float(np.sin(2 * np.pi * self.frequency * t + self.phase))


class TimeFunctions(Controller):
    def __init__(self, functions: Grid):
        self.functions = functions

    def get_functions(self) -> Grid:
        return self.functions


class CentralizedSensing(Controller):
    def __init__(self, body: Grid, function: TimedRealFunction):
        self.function = function
        self.n_in, self.n_out = self.n_of_inputs(body), self.n_of_outputs(body)

    @staticmethod
    def n_of_inputs(body: Grid) -> int:
        return sum(v.sensor_dim() for v in body.values() if v is not None)

    @staticmethod
    def n_of_outputs(body: Grid) -> int:
        return body.count(lambda v: v is not None)


class DistributedSensing(Controller):
    N_DIRS = 4

    def __init__(self, body: Grid, signals: int):
        self.signals = signals
        self.functions: Grid = Grid.create(body)

    @staticmethod
    def n_of_inputs(voxel: Voxel, signals: int) -> int:
        return voxel.sensor_dim() + signals * DistributedSensing.N_DIRS

    @staticmethod
    def n_of_outputs(voxel: Voxel, signals: int) -> int:
        return 1 + signals * DistributedSensing.N_DIRS

    def get_functions(self) -> Grid:
        return self.functions


@dataclass
class Robot:
    controller: Controller
    voxels: Grid

    def get_voxels(self) -> Grid:
        return self.voxels

    def get_controller(self) -> Controller:
        return self.controller


class SerializationUtils:
    @staticmethod
    def clone(obj):
        return copy.deepcopy(obj)


class RobotUtils:
    @staticmethod
    def build_shape(name: str) -> Grid:
        m = _ffi.shape_mask(name)
        return Grid(m.shape[1], m.shape[0], lambda x, y: bool(m[y, x]))

    @staticmethod
    def build_sensorizing_function(name: str) -> Callable[[Grid], Grid]:
        """RobotUtils.buildSensorizingFunction(name) (Starter.java:404-405): `uniform-<s1+s2..>-<sigma>`, `spinedTouch-<t|f>-<t|f>-<sigma>`,
        `spinedTouchSighted-...`, `empty`; the lists may differ between voxels (px, py, l1, l5, the spined configurations)."""
        def fn(body: Grid) -> Grid:
            if not name:
                return Grid(body.w, body.h, lambda x, y: Voxel([]) if body.get(x, y) else None)
            sensors, off = _ffi.sensor_config_voxels(name, body.mask())
            vox = iter(range(len(off) - 1))

            def make(x, y):
                if not body.get(x, y):
                    return None
                v = next(vox)
                return Voxel(list(sensors[off[v]:off[v + 1]]))

            return Grid(body.w, body.h, make)  # Grid's filler runs in grid order (y outer, x inner): the order of `off`

        return fn

    @staticmethod
    def build_robot_transformation(name: str, random: "_ffi.JavaRandomState | None" = None) -> Callable[[Robot], Robot]:
        """RobotUtils.buildRobotTransformation(name, externalRandom) [U] (call sites Starter.java:126,428): `identity`, or
        `broken-<ratio>-<seed|rnd>`: every voxel is replaced, with probability `ratio`, by a BreakableVoxel with a frozen actuator;
        the Random is created ONCE here (seed) or is the caller's (`rnd`), so consecutive robots continue one stream.
        `breakable-...` (malfunctions triggered during the episode) raises VSR_EUNSUPPORTED."""
        t = _ffi.robot_transformation(name)
        if t[0] == "identity":
            return lambda r: r
        _, ratio, seed = t
        rnd = _ffi.JavaRandomState(seed) if seed is not None else random
        if rnd is None:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, "transformation with seed `rnd` needs the run's Random")

        def transform(robot: Robot) -> Robot:
            def f(v):
                if v is None:
                    return None
                if rnd.next_double() > ratio:
                    return v
                return BreakableVoxel(list(v.sensors), seed=rnd.next_int())
            body = SerializationUtils.clone(robot.get_voxels())
            return Robot(robot.get_controller(), Grid.create(body, f))

        return transform

    @staticmethod
    def frozen_mask(robot: Robot) -> int:
        """bit v (voxel order) = the actuator of voxel v is frozen: what vsr_eval_ex takes as frozen_masks."""
        m, i = 0, 0
        for v in robot.get_voxels().values():
            if v is None:
                continue
            if getattr(v, "actuator_frozen", False):
                m |= 1 << i
            i += 1
        return m
