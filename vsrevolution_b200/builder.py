"""Host-side mirror of the reference's genotype -> robot mappers that sit on the hot path.

Mirrors, with the same names, parameters, dimension checks and error text:
  NamedProvider              builder/NamedProvider.java:13-65
  PrototypedFunctionBuilder  builder/PrototypedFunctionBuilder.java:11-104 (buildFor / exampleFor / compose)
  MLP                        builder/function/MLP.java:16-95
  BrainCentralized           builder/robot/BrainCentralized.java:19-59
  BrainPhaseValues           builder/robot/BrainPhaseValues.java:20-62
  BrainHomoDistributed       builder/robot/BrainHomoDistributed.java:22-98
  build_mapper               locomotion/Starter.java:97-116 (split on "<", compose left to right)
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import _ffi
from .robot import (CentralizedSensing, DistributedSensing, Grid, MultiLayerPerceptron, PhaseSin, RealFunction, Robot,
                    SerializationUtils, TimedRealFunction)

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


MAPPER_PROVIDER = NamedProvider.of({
    "brainCentralized": BrainCentralized(),
    "brainPhaseVals": BrainPhaseValues(),
    "brainHomoDist": BrainHomoDistributed(),
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
           "BrainHomoDistributed", "build_mapper", "SerializationUtils"]
