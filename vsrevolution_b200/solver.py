"""Host-side mirror of the solver the hot path sits in: `numGA` = `new DoublesStandard(0.75, 0.05, 3, 0.35)`
(locomotion/Starter.java:178; builder/solver/DoublesStandard.java:41-71), run as a device-resident generation loop
(include/vsr_ga.h): genomes never leave HBM, one episode-kernel launch per generation, only the best individual (and,
on request, the population) comes back.

    solver = DoublesStandard(0.75, 0.05, 3, 0.35).build({"nPop": "4096", "nEval": "40960"}).build(evaluator)
    best_genome, best_fitness = solver.solve(listener=print)

Multi-GPU: one process per GPU, GA state replicated (same seed -> identical offspring on every rank), each
rank evaluates its contiguous shard and the ranks all-gather 8 bytes per robot per generation (`torch.distributed`).
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from . import _ffi
from .locomotion import PopulationEvaluator


class _DeviceDoubles:
    """A device buffer of float64 exposed through __cuda_array_interface__, so torch can alias it without a copy."""

    def __init__(self, ptr: int, n: int, stream: int = 0):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2, "strides": None}


class GenerationLoop:
    """StandardEvolver(..., overlapping=true, remap) restated on the device; stop condition nOfFitnessEvaluations(nEval)."""

    def __init__(self, evaluator: PopulationEvaluator, n_pop: int, n_eval: int, x_over_prob: float = 0.75,
                 tournament_rate: float = 0.05, min_n_tournament: int = 3, sigma_mut: float = 0.35, remap: bool = False,
                 seed: int = 0, process_group=None, es: dict | None = None):
        self.evaluator = evaluator
        self.n_pop, self.n_eval = n_pop, n_eval
        self.remap, self.es = remap, es
        if not getattr(evaluator, "direct", True):
            raise _ffi.VsrError(_ffi.VSR_EUNSUPPORTED, "the device-resident loop needs a mapper whose genome is the controller "
                                "parameter vector; host-transforming mappers go through PopulationEvaluator.evaluate")
        # Per-evaluation inputs -- `-rnd` terrain seeds (Starter.java:89-93) and robot-transformation draws (Starter.java:427-429) --
        # are drawn on the host for every batch the loop evaluates and staged with vsr_stage_eval_opts before the batch is enqueued
        # (in batch order: with `remap` the re-evaluated parents come first).  Every rank of a multi-process run draws the same values.
        self._per_eval = evaluator.per_evaluation_inputs if getattr(evaluator, "has_per_evaluation_inputs", False) else None
        if es is not None:  # SimpleEvolutionaryStrategy(..., nOfParents, nOfElites, sigma, remap)
            self.ga = _ffi.GeneticAlgorithm(evaluator.ctx, n_pop, seed=seed, remap=remap, solver=_ffi.SOLVER_ES,
                                            sigma_mut=es["sigma"], n_parents=es["n_parents"], n_elites=es["n_elites"])
        else:
            self.ga = _ffi.GeneticAlgorithm(evaluator.ctx, n_pop, seed=seed, remap=remap, xover_prob=x_over_prob,
                                            tournament_rate=tournament_rate, min_tournament=min_n_tournament, sigma_mut=sigma_mut)
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
            self.ga.set_shard(dist.get_rank(process_group), self.world)

    # one exchange: every rank contributes its shard of the pending fitness vector
    def _exchange(self):
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        ptr, count, lo, hi, _ = self.ga.pending()
        per = (count + self.world - 1) // self.world
        self.ga.sync()  # shard evaluated
        whole = torch.as_tensor(_DeviceDoubles(ptr, per * self.world), device="cuda")  # padded to world * per (<= 7 spare slots)
        mine = whole[dist.get_rank(self.pg) * per:(dist.get_rank(self.pg) + 1) * per].clone()
        dist.all_gather_into_tensor(whole, mine, group=self.pg)
        torch.cuda.current_stream().synchronize()

    def _stage(self, count: int):
        if self._per_eval is not None:
            frozen, terrains = self._per_eval(count)
            self.evaluator.ctx.stage_eval_opts(count, frozen, terrains)

    def init(self):
        self._stage(self.n_pop)
        self.ga.init_begin()
        self._exchange()
        self.ga.init_end()

    def step(self):
        if self.es is None:   # numGA: offspring, preceded by the re-evaluated parents with remap
            self._stage(2 * self.n_pop if self.remap else self.n_pop)
        else:                 # ES: offspring only; the elites are re-evaluated too with remap
            self._stage(self.n_pop if self.remap else self.n_pop - self.es["n_elites"])
        self.ga.step_begin()
        self._exchange()
        self.ga.step_end()

    def solve(self, listener: Callable[[dict], None] | None = None):
        self.init()
        while self.ga.evaluations < self.n_eval:
            if listener is not None:
                listener({"iterations": self.ga.generation, "evals": self.ga.evaluations, "best.fitness": self.ga.best_fitness()})
            self.step()
        overflows = self.evaluator.ctx.contact_overflows()
        if overflows:  # never silent: some fitness values of this run came from a warp whose contact pool was full
            import warnings
            warnings.warn(f"{overflows} warp launches overflowed the contact pool; fitness values of those robots are unreliable")
        return self.ga.best()


class _SolverBuilder:
    def __init__(self, outer: "DoublesStandard", n_pop: int, n_eval: int, remap: bool):
        self.outer, self.n_pop, self.n_eval, self.remap = outer, n_pop, n_eval, remap

    def build(self, evaluator: PopulationEvaluator, seed: int = 0, process_group=None) -> GenerationLoop:
        o = self.outer
        return GenerationLoop(evaluator, self.n_pop, self.n_eval, o.x_over_prob, o.tournament_rate, o.min_n_tournament, o.sigma_mut,
                              self.remap, seed, process_group)


class DoublesStandard:
    """builder/solver/DoublesStandard.java: (xOverProb, tournamentRate, minNTournament, sigmaMut); `build(params)` reads
    nPop, nEval, diversity, remap (:42-45)."""

    def __init__(self, x_over_prob: float, tournament_rate: float, min_n_tournament: int, sigma_mut: float):
        self.x_over_prob, self.tournament_rate = x_over_prob, tournament_rate
        self.min_n_tournament, self.sigma_mut = min_n_tournament, sigma_mut

    def build(self, params: dict) -> _SolverBuilder:
        try:
            n_pop, n_eval = int(params["nPop"]), int(params["nEval"])
        except KeyError as e:  # Integer.parseInt(null) -> NumberFormatException in the reference
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, f"missing solver parameter {e.args[0]}")
        # `diversity=true` builds a StandardWithEnforcedDiversityEvolver (DoublesStandard.java:72-87): an offspring whose genotype
        # already exists in the population is re-drawn, at most 100 times.  Every offspring of this solver went through
        # GaussianMutation (both operators end with it, :54-57), so a duplicate of an existing List<Double> has probability zero and
        # the re-draw never triggers: the enforced-diversity loop is the standard loop, and the same device loop serves both.
        return _SolverBuilder(self, n_pop, n_eval, str(params.get("remap", "false")).lower() == "true")


class _ESBuilder:
    def __init__(self, outer: "SimpleES", n_pop: int, n_eval: int, remap: bool):
        self.outer, self.n_pop, self.n_eval, self.remap = outer, n_pop, n_eval, remap

    def build(self, evaluator: PopulationEvaluator, seed: int = 0, process_group=None) -> GenerationLoop:
        n_parents = int(np.floor(self.n_pop * self.outer.parents_rate + 0.5))  # (int) Math.round(nPop * parentsRate)
        return GenerationLoop(evaluator, self.n_pop, self.n_eval, remap=self.remap, seed=seed, process_group=process_group,
                              es={"sigma": self.outer.sigma, "n_parents": n_parents, "n_elites": 1})


class SimpleES:
    """builder/solver/SimpleES.java: (sigma, parentsRate); `build(params)` reads nPop, nEval, remap (:31-33) and builds a JGEA
    SimpleEvolutionaryStrategy with nOfParents = round(nPop * parentsRate), nOfElites = 1 (:41-49)."""

    def __init__(self, sigma: float, parents_rate: float):
        self.sigma, self.parents_rate = sigma, parents_rate

    def build(self, params: dict) -> _ESBuilder:
        try:
            n_pop, n_eval = int(params["nPop"]), int(params["nEval"])
        except KeyError as e:
            raise _ffi.VsrIllegalArgument(_ffi.VSR_EINVAL, f"missing solver parameter {e.args[0]}")
        return _ESBuilder(self, n_pop, n_eval, str(params.get("remap", "false")).lower() == "true")


def es() -> SimpleES:
    """`ES` (locomotion/Starter.java:179)."""
    return SimpleES(0.35, 0.4)


def num_ga() -> DoublesStandard:
    """`numGA` (locomotion/Starter.java:178)."""
    return DoublesStandard(0.75, 0.05, 3, 0.35)


def tournament_size(n_pop: int, tournament_rate: float = 0.05, min_n_tournament: int = 3) -> int:
    d = _ffi.GaDesc()
    _ffi.check(_ffi.lib().vsr_ga_default(d))
    d.n_pop, d.tournament_rate, d.min_tournament = n_pop, tournament_rate, min_n_tournament
    return _ffi.check(_ffi.lib().vsr_ga_tournament_size(d))
