"""Host-side mirror of the reference's genotype -> robot mappers that sit on the hot path.

Mirrors, with the same names, parameters, dimension checks and error text:
  NamedProvider              builder/NamedProvider.java:13-65
  PrototypedFunctionBuilder  builder/PrototypedFunctionBuilder.java:11-104 (buildFor / exampleFor / compose)
  MLP                        builder/function/MLP.java:16-95
  BrainCentralized           builder/robot/BrainCentralized.java:19-59
  BrainPhaseValues           builder/robot/BrainPhaseValues.java:20-62
  BrainHomoDistributed       builder/robot/BrainHomoDistributed.java:22-98
  BrainHeteroDistributed     builder/robot/BrainHeteroDistributed.java:17-102
  BrainPhaseFunction         builder/robot/BrainPhaseFunction.java:20-58
  BodyBrainSinusoidal        builder/robot/BodyBrainSinusoidal.java:20-140
  FunctionsGrid              builder/misc/FunctionsGrid.java:15-64
  DirectNumbersGrid          builder/misc/DirectNumbersGrid.java:14-55
  FunctionNumbersGrid        builder/misc/FunctionNumbersGrid.java:14-66
  build_mapper               locomotion/Starter.java:97-116 (split on "<", compose left to right)
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import _ffi
from .robot import (CentralizedSensing, DistributedSensing, Grid, MultiLayerPerceptron, PhaseSin, RealFunction, Robot,
                    SerializationUtils, Sinusoid, TimedRealFunction, TimeFunctions)

TOKEN_SEPARATOR = ";"
PARAM_VALUE_SEPARATOR = "="
NAME_KEY = "NAME"
MAPPER_PIPE_CHAR = "<"


def illegal(msg: str) -> _ffi.VsrIllegalArgument:
    return _ffi.VsrIllegalArgument(_ffi.VSR_EDIM, msg)


class NamedProvider:
    """`name;k=v;...` -> builder (builder/NamedProvider.java:46-60)."""

    def __init__(self, providers: dict):
        self.providers = providers

    @staticmethod
    def of(providers: dict) -> "NamedProvider":
        return NamedProvider(providers)

    @staticmethod
    def parse(string: str) -> dict:
        params = {}
        for piece in string.split(TOKEN_SEPARATOR):
            kv = piece.split(PARAM_VALUE_SEPARATOR)
            if len(kv) == 1:
                params[NAME_KEY] = kv[0]
            else:
                params[kv[0]] = kv[1]
        return params

    def build(self, string: str):
        params = self.parse(string)
        provider = self.providers.get(params.get(NAME_KEY))
        return None if provider is None else provider.build(params)


class PrototypedFunctionBuilder:
    def __init__(self, build_for: Callable, example_for: Callable):
        self._build_for, self._example_for = build_for, example_for

    def build_for(self, target):
        return self._build_for(target)

    def example_for(self, target):
        return self._example_for(target)

    def compose(self, other: "PrototypedFunctionBuilder") -> "PrototypedFunctionBuilder":
        this = self

        def build_for(c):
            outer = this.build_for(c)
            inner = other.build_for(this.example_for(c))
            return lambda a: outer(inner(a))

        return PrototypedFunctionBuilder(build_for, lambda c: other.example_for(this.example_for(c)))


class MLP:
    def build(self, params: dict) -> PrototypedFunctionBuilder:
        ratio = float(params.get("r", "0.65"))
        n_inner = int(params.get("nIL", "1"))

        def build_for(function: TimedRealFunction):
            def f(values):
                n_in, n_out = function.get_input_dimension(), function.get_output_dimension()
                inner = MLP.inner_neurons(n_in, n_out, ratio, n_inner)
                n_w = MultiLayerPerceptron.count_weights(n_in, inner, n_out)
                if n_w != len(values):
                    raise illegal(f"Wrong number of values for weights: {n_w} expected, {len(values)} found")
                return MultiLayerPerceptron(MultiLayerPerceptron.TANH, n_in, inner, n_out, np.asarray(values, dtype=np.float64))

            return f

        def example_for(function: TimedRealFunction):
            n_in, n_out = function.get_input_dimension(), function.get_output_dimension()
            return [0.0] * MultiLayerPerceptron.count_weights(n_in, MLP.inner_neurons(n_in, n_out, ratio, n_inner), n_out)

        return PrototypedFunctionBuilder(build_for, example_for)

    @staticmethod
    def inner_neurons(n_in: int, n_out: int, ratio: float, n_inner: int) -> list[int]:
        return _ffi.mlp_inner_neurons(n_in, n_out, ratio, n_inner)


class BrainCentralized:
    def build(self, params: dict) -> PrototypedFunctionBuilder:
        def build_for(robot: Robot):
            body = robot.get_voxels()

            def f(function: TimedRealFunction):
                if function.get_input_dimension() != CentralizedSensing.n_of_inputs(body):
                    raise illegal(f"Wrong number of function input args: {CentralizedSensing.n_of_inputs(body)} expected, "
                                  f"{function.get_input_dimension()} found")
                if function.get_output_dimension() != CentralizedSensing.n_of_outputs(body):
                    raise illegal(f"Wrong number of function output args: {CentralizedSensing.n_of_outputs(body)} expected, "
                                  f"{function.get_output_dimension()} found")
                return Robot(CentralizedSensing(body, function), body)

            return f

        def example_for(robot: Robot):
            body = robot.get_voxels()
            return RealFunction.build(None, CentralizedSensing.n_of_inputs(body), CentralizedSensing.n_of_outputs(body))

        return PrototypedFunctionBuilder(build_for, example_for)


class BrainPhaseValues:
    def build(self, params: dict) -> PrototypedFunctionBuilder:
        frequency = float(params.get("f", "1"))
        amplitude = float(params.get("a", "1"))

        def build_for(robot: Robot):
            body = robot.get_voxels()
            n_vox = body.count(lambda v: v is not None)

            def f(values):
                if n_vox != len(values):
                    raise illegal(f"Wrong number of values: {n_vox} expected, {len(values)} found")
                c = 0
                phases = Grid.create(body)
                for entry in body:
                    if entry.value is not None:
                        phases.set(entry.key.x, entry.key.y, float(values[c]))
                        c += 1
                return Robot(PhaseSin(frequency, amplitude, phases), body)

            return f

        return PrototypedFunctionBuilder(build_for, lambda robot: [0.0] * robot.get_voxels().count(lambda v: v is not None))


class BrainHomoDistributed:
    @staticmethod
    def _io_dim(robot: Robot, signals: int):
        body = robot.get_voxels()
        first = next((v for v in body.values() if v is not None), None)
        if first is None:
            raise illegal("Target robot has no voxels")
        n_in = DistributedSensing.n_of_inputs(first, signals)
        n_out = DistributedSensing.n_of_outputs(first, signals)
        wrong = [e for e in body if e.value is not None and DistributedSensing.n_of_inputs(e.value, signals) != n_in]
        if wrong:
            raise illegal(f"Cannot build BrainHomoDistributed robot mapper for this body: all voxels should have {n_in} "
                          "inputs, but voxels at positions " + ",".join(f"({e.key.x},{e.key.y})" for e in wrong) + " have " +
                          ",".join(str(DistributedSensing.n_of_inputs(e.value, signals)) for e in wrong))
        return n_in, n_out

    def build(self, params: dict) -> PrototypedFunctionBuilder:
        signals = int(params.get("s", "1"))

        def build_for(robot: Robot):
            n_in, n_out = self._io_dim(robot, signals)
            body = robot.get_voxels()

            def f(function: TimedRealFunction):
                if function.get_input_dimension() != n_in:
                    raise illegal(f"Wrong number of function input args: {n_in} expected, {function.get_input_dimension()} found")
                if function.get_output_dimension() != n_out:
                    raise illegal(f"Wrong number of function output args: {n_out} expected, {function.get_output_dimension()} found")
                controller = DistributedSensing(body, signals)
                for entry in body:
                    if entry.value is not None:
                        controller.get_functions().set(entry.key.x, entry.key.y, function)  # homogeneous: shared, immutable
                return Robot(controller, body)

            return f

        def example_for(robot: Robot):
            n_in, n_out = self._io_dim(robot, signals)
            return RealFunction.build(None, n_in, n_out)

        return PrototypedFunctionBuilder(build_for, example_for)


class BrainHeteroDistributed:
    """`brainHeteroDist;s=<signals>`: DistributedSensing with one function per voxel (BrainHeteroDistributed.java:36-83)."""

    @staticmethod
    def _io_dims(robot: Robot, signals: int) -> Grid:
        return Grid.create(robot.get_voxels(), lambda v: None if v is None else (
            DistributedSensing.n_of_inputs(v, signals), DistributedSensing.n_of_outputs(v, signals)))

    def build(self, params: dict) -> PrototypedFunctionBuilder:
        signals = int(params.get("s", "1"))

        def build_for(robot: Robot):
            dims = self._io_dims(robot, signals)
            body = robot.get_voxels()

            def f(functions: Grid):
                if dims.w != functions.w or dims.h != functions.h:
                    raise illegal(f"Wrong size of functions grid: {dims.w}x{dims.h} expected, {functions.w}x{functions.h} found")
                for entry in dims:
                    if entry.value is None:
                        continue
                    fn = functions.get(entry.key.x, entry.key.y)
                    if fn.get_input_dimension() != entry.value[0]:
                        raise illegal(f"Wrong number of function input args at ({entry.key.x},{entry.key.y}): "
                                      f"{entry.value[0]} expected, {fn.get_input_dimension()} found")
                    if fn.get_output_dimension() != entry.value[1]:
                        raise illegal(f"Wrong number of function output args at ({entry.key.x},{entry.key.y}): "
                                      f"{entry.value[1]} expected, {fn.get_output_dimension()} found")
                controller = DistributedSensing(body, signals)
                for entry in body:
                    if entry.value is not None:
                        controller.get_functions().set(entry.key.x, entry.key.y, functions.get(entry.key.x, entry.key.y))
                return Robot(controller, body)

            return f

        def example_for(robot: Robot):
            return Grid.create(self._io_dims(robot, signals), lambda d: None if d is None else RealFunction.build(None, d[0], d[1]))

        return PrototypedFunctionBuilder(build_for, example_for)


class FunctionsGrid:
    """`funsGrid`: slices one flat genome into the voxels' functions, non-null cells in Grid iteration order
    (FunctionsGrid.java:31-46)."""

    def __init__(self, item_builder: PrototypedFunctionBuilder):
        self.item_builder = item_builder

    def build(self, params: dict) -> PrototypedFunctionBuilder:
        item = self.item_builder

        def build_for(target_functions: Grid):
            def f(values):
                functions = Grid.create(target_functions)
                c = 0
                for entry in target_functions:
                    if entry.value is None:
                        continue
                    size = len(item.example_for(entry.value))
                    functions.set(entry.key.x, entry.key.y, item.build_for(entry.value)(values[c:c + size]))
                    c += size
                return functions

            return f

        def example_for(functions: Grid):
            return [0.0] * sum(len(item.example_for(fn)) for fn in functions.values() if fn is not None)

        return PrototypedFunctionBuilder(build_for, example_for)


class BrainPhaseFunction:
    """`brainPhaseFun;f;a`: the phase of voxel (x, y) is function(x / w, y / w) -- the reference divides BOTH coordinates by the
    width (BrainPhaseFunction.java:38-42) and fills every cell of the grid, voxel or not."""

    def build(self, params: dict) -> PrototypedFunctionBuilder:
        frequency = float(params.get("f", "1"))
        amplitude = float(params.get("a", "1"))

        def build_for(robot: Robot):
            body = robot.get_voxels()

            def f(function: TimedRealFunction):
                if function.get_input_dimension() != 2:
                    raise illegal(f"Wrong number of function input args: 2 expected, {function.get_input_dimension()} found")
                phases = Grid(body.w, body.h, lambda x, y: float(function.apply(0.0, [x / body.w, y / body.w])[0]))
                return Robot(PhaseSin(frequency, amplitude, phases), body)

            return f

        return PrototypedFunctionBuilder(build_for, lambda robot: RealFunction.build(None, 2, 1))


def percentile_legacy(values, p: float) -> float:
    """commons-math3 `new Percentile().evaluate(values, p)` (default LEGACY estimation, p in (0, 100]) as
    BodyBrainSinusoidal.java:74-77 calls it: pos = p (n + 1) / 100 on the sorted values, linear interpolation."""
    if not (0 < p <= 100):
        raise illegal(f"out of bounds quantile value: {p}, must be in (0, 100]")
    xs = sorted(float(v) for v in values)
    n = len(xs)
    if n == 0:
        return float("nan")
    if n == 1:
        return xs[0]
    pos = p * (n + 1) / 100.0
    fpos = np.floor(pos)
    if pos < 1:
        return xs[0]
    if pos >= n:
        return xs[-1]
    lower, upper = xs[int(fpos) - 1], xs[int(fpos)]
    return lower + (pos - fpos) * (upper - lower)


def grid_largest_connected(grid: Grid, pred) -> Grid:
    """2dhmsr Utils.gridLargestConnected [U]: keep the largest 4-connected set of cells satisfying pred (first found wins ties)."""
    seen, best = set(), []
    for e in grid:
        k = (e.key.x, e.key.y)
        if k in seen or not pred(e.value):
            continue
        comp, stack = [], [k]
        seen.add(k)
        while stack:
            x, y = stack.pop()
            comp.append((x, y))
            for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                q = (x + dx, y + dy)
                if 0 <= q[0] < grid.w and 0 <= q[1] < grid.h and q not in seen and pred(grid.get(*q)):
                    seen.add(q)
                    stack.append(q)
        if len(comp) > len(best):
            best = comp
    keep = set(best)
    return Grid(grid.w, grid.h, lambda x, y: grid.get(x, y) if (x, y) in keep else None)


def crop_grid(grid: Grid, pred) -> Grid:
    """2dhmsr Utils.cropGrid [U]: the bounding box of the cells satisfying pred."""
    cells = [(e.key.x, e.key.y) for e in grid if pred(e.value)]
    if not cells:
        return Grid(0, 0)
    x0, x1 = min(c[0] for c in cells), max(c[0] for c in cells)
    y0, y1 = min(c[1] for c in cells), max(c[1] for c in cells)
    return Grid(x1 - x0 + 1, y1 - y0 + 1, lambda x, y: grid.get(x + x0, y + y0))


class BodyBrainSinusoidal:
    """`bodyBrainSin;minF;maxF;p`: a grid of (body value, [frequency], [phase], [amplitude]) numbers -> body = the largest
    connected set of cells whose body value reaches the p-th percentile, controller = TimeFunctions of per-voxel sinusoids
    (BodyBrainSinusoidal.java:36-124).  With the Starter's components {PHASE, FREQUENCY} (Starter.java:207-210) every cell
    holds 3 numbers in the order (body, frequency, phase)."""
    FREQUENCY, PHASE, AMPLITUDE = "FREQUENCY", "PHASE", "AMPLITUDE"

    def __init__(self, components=(PHASE, FREQUENCY)):
        self.components = set(components)

    @staticmethod
    def _clip(v: float) -> float:
        return min(max(v, -1.0), 1.0)

    def build(self, params: dict) -> PrototypedFunctionBuilder:
        min_f = float(params.get("minF", "0.5"))
        max_f = float(params.get("maxF", "2"))
        percentile = float(params.get("p", "0.5"))
        n_vals = 1 + len(self.components)

        def build_for(robot: Robot):
            prototype = next((v for v in robot.get_voxels().values() if v is not None), None)
            if prototype is None:
                raise illegal("Target robot has no valid voxels")
            tw, th = robot.get_voxels().w, robot.get_voxels().h

            def f(grid: Grid):
                if grid.w != tw or grid.h != th:
                    raise illegal(f"Wrong grid size: {tw}x{th} expected, {grid.w}x{grid.h} found")
                wrong = next((e for e in grid if len(e.value) != n_vals), None)
                if wrong is not None:
                    raise illegal(f"Wrong number of values in grid at {wrong.key.x},{wrong.key.y}: {n_vals} expected, "
                                  f"{len(wrong.value)} found")
                threshold = percentile_legacy([v[0] for v in grid.values()], percentile) if percentile > 0 else 0.0
                cropped = crop_grid(grid_largest_connected(
                    Grid.create(grid, lambda v: v if v[0] >= threshold else None), lambda v: v is not None), lambda v: v is not None)
                body = Grid.create(cropped, lambda v: SerializationUtils.clone(prototype) if v is not None else None)
                if body.count(lambda v: v is not None) == 0:
                    body = Grid(1, 1, SerializationUtils.clone(prototype))

                def sinusoid(x, y):
                    vs = cropped.get(x, y)
                    if vs is None:
                        vs = [0.0] * 4
                    c = 1
                    if self.FREQUENCY in self.components:
                        freq = min_f + (max_f - min_f) * (self._clip(vs[c]) + 1.0) / 2.0
                        c += 1
                    else:
                        freq = (min_f + max_f) / 2.0
                    if self.PHASE in self.components:
                        phase = np.pi * (self._clip(vs[c]) + 1.0) / 2.0
                        c += 1
                    else:
                        phase = 0.0
                    if self.AMPLITUDE in self.components:
                        amplitude = (self._clip(vs[c]) + 1.0) / 2.0
                        c += 1
                    else:
                        amplitude = 1.0
                    return Sinusoid(float(freq), float(phase), float(amplitude))

                return Robot(TimeFunctions(Grid(body.w, body.h, sinusoid)), body)

            return f

        def example_for(robot: Robot):
            if not any(v is not None for v in robot.get_voxels().values()):
                raise illegal("Target robot has no valid voxels")
            return Grid(robot.get_voxels().w, robot.get_voxels().h, lambda x, y: [0.0] * n_vals)

        return PrototypedFunctionBuilder(build_for, example_for)


class DirectNumbersGrid:
    """`dirNumGrid`: flat values -> Grid<double[]>, filled x OUTER, y inner (DirectNumbersGrid.java:32-38) -- column-major,
    unlike Grid's own row-major iteration."""

    def build(self, params: dict) -> PrototypedFunctionBuilder:
        def build_for(grid: Grid):
            expected = sum(len(v) for v in grid.values())

            def f(values):
                if expected != len(values):
                    raise illegal(f"Wrong number of values: {expected} expected, {len(values)} found")
                c = 0
                output = Grid.create(grid)
                for x in range(output.w):
                    for y in range(output.h):
                        n = len(grid.get(x, y))
                        output.set(x, y, [float(v) for v in values[c:c + n]])
                        c += n
                return output

            return f

        return PrototypedFunctionBuilder(build_for, lambda grid: [0.0] * sum(len(v) for v in grid.values()))


class FunctionNumbersGrid:
    """`funNumGrid`: cell (x, y) = function(x / (w - 1), y / (h - 1)) (FunctionNumbersGrid.java:33-58)."""

    @staticmethod
    def _target_length(grid: Grid) -> int:
        lengths = sorted({len(v) for v in grid.values()})
        if len(lengths) > 1:
            raise illegal(f"Target grid elements must have all the same size, found {len(lengths)} distinct sizes {lengths}")
        return lengths[0]

    def build(self, params: dict) -> PrototypedFunctionBuilder:
        def build_for(grid: Grid):
            n = self._target_length(grid)

            def f(function: TimedRealFunction):
                if function.get_input_dimension() != 2 or function.get_output_dimension() != n:
                    raise illegal(f"Wrong number of body function args: 2->{n} expected, "
                                  f"{function.get_input_dimension()}->{function.get_output_dimension()} found")
                w, h = float(grid.w), float(grid.h)
                with np.errstate(divide="ignore", invalid="ignore"):  # a 1-wide grid divides by zero, as the Java does
                    return Grid(grid.w, grid.h, lambda x, y: [float(v) for v in function.apply(
                        [float(np.float64(x) / np.float64(w - 1.0)), float(np.float64(y) / np.float64(h - 1.0))])])

            return f

        return PrototypedFunctionBuilder(build_for, lambda grid: RealFunction.build(None, 2, self._target_length(grid)))


MAPPER_PROVIDER = NamedProvider.of({
    "brainCentralized": BrainCentralized(),
    "brainPhaseVals": BrainPhaseValues(),
    "brainHomoDist": BrainHomoDistributed(),
    "brainHeteroDist": BrainHeteroDistributed(),
    "brainPhaseFun": BrainPhaseFunction(),
    "bodyBrainSin": BodyBrainSinusoidal((BodyBrainSinusoidal.PHASE, BodyBrainSinusoidal.FREQUENCY)),  # Starter.java:207-210
    "dirNumGrid": DirectNumbersGrid(),
    "funNumGrid": FunctionNumbersGrid(),
    "funsGrid": FunctionsGrid(MLP().build({"r": "0.65", "nIL": "1"})),  # Starter.java:232-238
    "mlp": MLP(),
})


def build_mapper(robot_mapper_name: str, provider: Optional[NamedProvider] = None) -> PrototypedFunctionBuilder:
    """locomotion/Starter.java:105-113: first piece outermost, composed left to right."""
    provider = provider or MAPPER_PROVIDER
    mapper = None
    for piece in robot_mapper_name.split(MAPPER_PIPE_CHAR):
        b = provider.build(piece)
        if b is None:
            raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, f"mapper {piece} is outside the hot path (SURVEY.md 8)")
        mapper = b if mapper is None else mapper.compose(b)
    return mapper


__all__ = ["NamedProvider", "PrototypedFunctionBuilder", "MLP", "BrainCentralized", "BrainPhaseValues",
           "BrainHomoDistributed", "BrainHeteroDistributed", "BrainPhaseFunction", "BodyBrainSinusoidal", "FunctionsGrid",
           "DirectNumbersGrid", "FunctionNumbersGrid", "percentile_legacy", "build_mapper", "SerializationUtils"]
