"""GPU parity of the device-resident generation loop (include/vsr_ga.h) against oracle/ga.py, through the C ABI.

Integer decisions (parents, crossover masks via the genes they pick, survivors, ranks) must match index for index; genes
carry a Gaussian computed with device libm (log/sqrt/cos in double), so they are compared to 1e-12 absolute.  The fitness
values are whatever the episode kernel returns: the oracle GA is fed the device's own fitness vector, so the test is exact
and independent of the (chaotic) physics.
"""
import numpy as np
import pytest

from oracle import ga as OG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from vsrevolution_b200.locomotion import population_evaluator
    return population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 2.0, 0.5)


@pytest.mark.parametrize("remap", [False, True])
def test_generation_loop_matches_oracle_index_for_index(ev, remap):
    from vsrevolution_b200 import _ffi
    n, G = 96, ev.genome_len
    ga = _ffi.GeneticAlgorithm(ev.ctx, n, seed=11, remap=remap)
    og = OG.NumGA(n, G, seed=11, remap=remap)
    assert ga.tournament_size == og.nt == 5
    ga.init()
    pop, fit = ga.population()
    g0 = og.initial_genomes()
    f0 = ev.evaluate(g0)  # same kernel, same genomes: deterministic
    o = og.set_population(g0, f0)
    assert np.array_equal(ga.survivors(), o)
    assert np.array_equal(pop, og.pop) and np.array_equal(fit, og.fit)  # uniform init is exact arithmetic
    assert ga.generation == 0 and ga.evaluations == n
    for gen in range(1, 4):
        ga.step()
        off, par, off_fit = ga.offspring()
        opar = og.select(gen)
        assert np.array_equal(par, opar)
        ooff = og.vary(gen, opar)
        assert np.abs(off - ooff).max() < 1e-12
        # feed the oracle the device's genomes and fitness, so the merge is compared exactly
        if remap:
            allfit = ev.evaluate(np.concatenate([og.pop, off]))
            origin = og.merge(off, allfit[n:], allfit[:n])
        else:
            origin = og.merge(off, off_fit)
        pop, fit = ga.population()
        assert np.array_equal(ga.survivors(), origin)
        assert np.array_equal(pop, og.pop) and np.array_equal(fit, og.fit)
        assert np.all(np.diff(fit) <= 0)
        assert ga.generation == gen
    assert ga.evaluations == n + 3 * (2 * n if remap else n)
    bg, bf = ga.best()
    assert bf == fit[0] and np.array_equal(bg, pop[0])
    ga.close()


def test_split_begin_end_and_shards_equal_the_single_call(ev):
    """Two GA replicas evaluating half of every batch each (the N>1 path, exchanged by hand here) == one unsharded GA."""
    import torch
    from vsrevolution_b200 import _ffi
    from vsrevolution_b200.solver import _DeviceDoubles
    n = 50  # odd split: per = 25; with remap off counts are 50
    whole = _ffi.GeneticAlgorithm(ev.ctx, n, seed=4)
    whole.init()
    for _ in range(2):
        whole.step()
    ref_pop, ref_fit = whole.population()
    whole.close()
    # replicas share the context (serialised), each with its own GA state
    reps = [_ffi.GeneticAlgorithm(ev.ctx, n, seed=4) for _ in range(2)]
    for r, g in enumerate(reps):
        g.set_shard(r, 2)

    def exchange():
        views = []
        for g in reps:
            ptr, count, lo, hi, _ = g.pending()
            g.sync()
            views.append((torch.as_tensor(_DeviceDoubles(ptr, count), device="cuda"), lo, hi))
        for i, (v, lo, hi) in enumerate(views):
            for j, (w, _, _) in enumerate(views):
                if i != j:
                    w[lo:hi] = v[lo:hi]
        torch.cuda.synchronize()

    for g in reps:
        g.init_begin()
    exchange()
    for g in reps:
        g.init_end()
    for _ in range(2):
        for g in reps:
            g.step_begin()
        exchange()
        for g in reps:
            g.step_end()
    for g in reps:
        pop, fit = g.population()
        assert np.array_equal(pop, ref_pop) and np.array_equal(fit, ref_fit)
        g.close()


def test_set_population_sorts_and_nan_ranks_last(ev):
    from vsrevolution_b200 import _ffi
    n, G = 8, ev.genome_len
    ga = _ffi.GeneticAlgorithm(ev.ctx, n, seed=0)
    genomes = np.arange(n * G, dtype=np.float64).reshape(n, G)
    fit = np.array([0.5, np.nan, 2.0, 2.0, -1.0, np.inf, -np.inf, 0.0])
    ga.set_population(genomes, fit)
    pop, f = ga.population()
    o = OG.order(fit)
    assert np.array_equal(ga.survivors(), o)
    assert np.array_equal(pop, genomes[o])
    assert np.array_equal(f, fit[o], equal_nan=True)
    with pytest.raises(_ffi.VsrIllegalArgument):
        ga.set_population(genomes[:4], fit[:4])
    with pytest.raises(_ffi.VsrError):
        ga.step_end()
    ga.close()


def test_solver_mirror_improves_fitness(ev):
    from vsrevolution_b200.solver import num_ga
    log = []
    solver = num_ga().build({"nPop": "256", "nEval": str(256 * 6)}).build(ev, seed=2)
    genome, best = solver.solve(listener=log.append)
    assert solver.ga.evaluations == 256 * 6 and solver.ga.generation == 5
    assert [e["iterations"] for e in log] == [0, 1, 2, 3, 4]
    bests = [e["best.fitness"] for e in log] + [best]
    assert all(b >= a for a, b in zip(bests, bests[1:]))
    assert np.isclose(ev.evaluate(genome[None, :])[0], best, rtol=0, atol=0)  # deterministic re-evaluation of the best


def test_large_population_ranking_by_bitonic_sort(ev):
    """Merged sets above 8192 rows are ranked by a bitonic sort of (key, index) pairs instead of the all-pairs count: same strict
    total order (higher first, NaN as -inf, -0 == +0, ties towards the lower index), checked against oracle/ga.py:order."""
    from vsrevolution_b200 import _ffi
    n, G = 9001, ev.genome_len  # not a power of two: the sort pads to 16384
    rng = np.random.default_rng(8)
    fit = rng.normal(size=n).round(2)  # many ties
    fit[rng.integers(0, n, 40)] = np.nan
    fit[rng.integers(0, n, 40)] = 0.0
    fit[rng.integers(0, n, 40)] = -0.0
    fit[[5, 77]] = np.inf
    fit[[9, 8000]] = -np.inf
    genomes = rng.uniform(-1, 1, (n, G))
    ga = _ffi.GeneticAlgorithm(ev.ctx, n, seed=0)
    ga.set_population(genomes, fit)
    o = OG.order(fit)
    assert np.array_equal(ga.survivors(), o)
    pop, f = ga.population()
    assert np.array_equal(pop, genomes[o]) and np.array_equal(f, fit[o], equal_nan=True)
    ga.close()
    # and inside a generation (merged set of 2 x 5000 rows), index for index against the oracle
    n = 5000
    ga = _ffi.GeneticAlgorithm(ev.ctx, n, seed=21)
    og = OG.NumGA(n, G, seed=21)
    ga.init()
    g0 = og.initial_genomes()
    og.set_population(g0, ev.evaluate(g0))
    for gen in (1, 2):
        ga.step()
        off, par, off_fit = ga.offspring()
        assert np.array_equal(par, og.select(gen))
        origin = og.merge(off, off_fit)
        assert np.array_equal(ga.survivors(), origin)
        pop, f = ga.population()
        assert np.array_equal(f, og.fit) and np.array_equal(pop, og.pop)
    ga.close()


@pytest.mark.parametrize("remap", [False, True])
def test_es_generation_loop_matches_oracle(ev, remap):
    """`ES` = SimpleES(0.35, 0.4) (Starter.java:179; builder/solver/SimpleES.java:41-49) on the device vs oracle/ga.py: means bit
    for bit (same summation order), offspring genes to 1e-12 (device libm Gaussian), elites / ranks exact."""
    from vsrevolution_b200 import _ffi
    n, G = 50, ev.genome_len
    ga = _ffi.GeneticAlgorithm(ev.ctx, n, seed=5, remap=remap, solver=_ffi.SOLVER_ES)
    og = OG.SimpleES(n, G, seed=5, remap=remap)
    assert ga.desc.n_parents == og.n_parents == 20 and ga.desc.n_elites == 1 and ga.desc.sigma_mut == 0.35
    ga.init()
    g0 = og.initial_genomes()
    og.set_population(g0, ev.evaluate(g0))
    pop, fit = ga.population()
    assert np.array_equal(pop, og.pop) and np.array_equal(fit, og.fit)
    for gen in range(1, 4):
        ga.step()
        means, nxt = og.sample(gen)
        assert np.array_equal(ga.means(), means)
        off, par, off_fit = ga.offspring()        # rows of the new generation: the elite, then the offspring
        assert np.array_equal(off[:1], og.pop[:1]) and np.abs(off - nxt).max() < 1e-12 and (par == -1).all()
        if not remap:
            assert off_fit[0] == og.fit[0]        # the elite keeps its fitness
        o = OG.order(off_fit)
        og.pop, og.fit, og.gen = off[o], off_fit[o], gen
        pop, fit = ga.population()
        assert np.array_equal(ga.survivors(), o)
        assert np.array_equal(pop, og.pop) and np.array_equal(fit, og.fit) and np.all(np.diff(fit) <= 0)
    assert ga.evaluations == n + 3 * (n if remap else n - 1)
    if not remap:  # elitism: the best fitness never decreases
        assert fit[0] >= ev.evaluate(g0).max()
    ga.close()
    with pytest.raises(_ffi.VsrIllegalArgument):
        _ffi.GeneticAlgorithm(ev.ctx, n, solver=_ffi.SOLVER_ES, n_parents=n + 1)


def test_es_solver_mirror_improves(ev):
    from vsrevolution_b200 import solver as S
    loop = S.es().build({"nPop": "64", "nEval": "400"}).build(ev, seed=2)
    seen = []
    best_g, best_f = loop.solve(listener=lambda d: seen.append(d["best.fitness"]))
    assert len(seen) >= 5 and all(b >= a for a, b in zip(seen, seen[1:])) and best_f >= seen[0]
    assert loop.ga.evaluations >= 400 and len(best_g) == ev.genome_len


def test_generation_loop_draws_per_evaluation_terrains_and_transformations():
    """Starter.java:89-93 draws a terrain seed per evaluation for `-rnd` names and Starter.java:427-429 applies the robot
    transformation per evaluation: the device loop stages them per batch (vsr_stage_eval_opts) instead of silently evaluating every
    individual on seed 0 / untransformed.  The loop's fitness of an offspring == that robot evaluated alone on its
    terrain with its frozen mask."""
    from vsrevolution_b200 import _ffi
    from vsrevolution_b200.locomotion import population_evaluator
    from vsrevolution_b200.solver import num_ga
    kw = dict(transformation="broken-0.3-7")
    ev = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "hilly-1-10-rnd", 2.0, 0.5, **kw)
    twin = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "hilly-1-10-rnd", 2.0, 0.5, **kw)
    assert ev.has_per_evaluation_inputs
    n = 48
    loop = num_ga().build({"nPop": str(n), "nEval": str(3 * n)}).build(ev, seed=5)
    loop.init()
    frozen0, terrains0 = twin.per_evaluation_inputs(n)          # the same draws, replayed: generation 0
    loop.step()
    frozen1, terrains1 = twin.per_evaluation_inputs(n)          # generation 1 continues both streams
    assert frozen0.any() and not np.array_equal(frozen0, frozen1)
    assert not np.array_equal(terrains0[0][1], terrains1[0][1]) and not np.array_equal(terrains1[0][1], terrains1[1][1])
    g, _, f = loop.ga.offspring()
    ref = twin.ctx.eval_ex(g, frozen_masks=frozen1, terrains=terrains1)[0]
    assert np.array_equal(f, ref)
    plain = twin.ctx.eval(g)                                       # ... and NOT what the context's own terrain / intact robots give
    assert (plain != f).mean() > 0.9
    # ES stages n - n_elites entries per generation
    from vsrevolution_b200.solver import es
    loop2 = es().build({"nPop": "32", "nEval": "96"}).build(ev, seed=1)
    loop2.solve()
    assert np.isfinite(loop2.ga.best()[1])


def test_overflowing_robots_rank_last_in_the_generation_loop(monkeypatch):
    """A robot whose warp dropped a manifold gets NaN (vsr_kernel.cuh epilogue); NaN ranks last, so the loop cannot select on it."""
    from vsrevolution_b200 import _ffi
    from vsrevolution_b200.locomotion import population_evaluator
    monkeypatch.setenv("VSR_POOL_SLOTS", "8")
    small = population_evaluator("biped-4x3", "uniform-a-0", "brainCentralized<mlp;r=0.65;nIL=1", "flat", 1.5, 0.5)
    monkeypatch.delenv("VSR_POOL_SLOTS")
    ga = _ffi.GeneticAlgorithm(small.ctx, 48, seed=3)
    ga.init()
    ga.step()
    _, fit = ga.population()
    assert small.ctx.contact_overflows() > 0
    nan = np.isnan(fit)
    assert nan.any()
    first_nan = int(np.argmax(nan))
    assert nan[first_nan:].all() and not nan[:first_nan].any()              # sorted best-first, every NaN behind every number
    ga.close()
