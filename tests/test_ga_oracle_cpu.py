"""CPU checks of the generation-loop oracle (oracle/ga.py) and of the GA boundary (include/vsr_ga.h).

The reference holds no tests or vectors for its solver; what can be pinned is (1) Philox4x32-10 against the
Random123 known-answer vectors, (2) the integer rules read directly from builder/solver/DoublesStandard.java (tournament
size, operator weights), (3) the statistical semantics of every operator, (4) the exported symbols.
"""
import ctypes as C
import os
import re

import numpy as np

from oracle import ga as OG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0), (0, 0, 0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF,) * 4, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0xA4093822, 0x299F31D0), (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for key, ctr, exp in kat:
        out = OG.philox4x32(key, [np.array([c]) for c in ctr])
        assert tuple(int(x[0]) for x in out) == exp


def test_tournament_size_follows_the_reference_formula():
    # Math.max(minNTournament, (int) Math.ceil((double) nPop * tournamentRate))   DoublesStandard.java:65
    assert OG.tournament_size(100) == 5
    assert OG.tournament_size(10) == 3
    assert OG.tournament_size(4096) == 205
    assert OG.tournament_size(1 << 18) == 13108
    assert OG.tournament_size(41, 0.05, 3) == 3  # ceil(2.05) = 3
    assert OG.tournament_size(61, 0.05, 3) == 4


def test_order_is_a_strict_total_order():
    f = np.array([1.0, np.nan, 3.0, 3.0, -np.inf, 0.0, -0.0])
    assert list(OG.order(f)) == [2, 3, 0, 5, 6, 1, 4]


def test_operator_statistics():
    ga = OG.NumGA(2000, 40, seed=3)
    g0 = ga.initial_genomes()
    assert g0.min() >= -1 and g0.max() < 1 and abs(g0.mean()) < 0.02 and abs(g0.var() - 1 / 3) < 0.02
    ga.set_population(g0, np.arange(2000, 0, -1.0))  # already sorted best-first
    par = ga.select(1)
    assert abs((par[:, 1] >= 0).mean() - 0.75) < 0.03  # P(crossover) = xOverProb
    # min of nT uniform draws over [0, n): E = n / (nT + 1)
    assert ga.nt == 100 and abs(par[:, 0].mean() - 2000 / 101) < 2.0
    off = ga.vary(1, par)
    mut_only = par[:, 1] < 0
    d = off[mut_only] - ga.pop[par[mut_only, 0]]
    assert abs(d.mean()) < 0.01 and abs(d.std() - 0.35) < 0.01  # GaussianMutation(sigma) on every gene
    x = ~mut_only
    d1, d2 = off[x] - ga.pop[par[x, 0]], off[x] - ga.pop[par[x, 1]]
    from_p1 = np.abs(d1) < np.abs(d2)  # crude attribution; both parents contribute about half of the genes
    assert 0.4 < from_p1.mean() < 0.6
    # draws are pure functions of (seed, generation, individual): another instance reproduces them
    gb = OG.NumGA(2000, 40, seed=3)
    gb.set_population(g0, np.arange(2000, 0, -1.0))
    assert np.array_equal(gb.select(1), par) and np.array_equal(gb.vary(1, par), off)
    assert not np.array_equal(gb.select(2), par)


def test_overlapping_truncation_keeps_the_best_and_is_monotone():
    ga = OG.NumGA(64, 10, seed=1)
    fn = lambda G: G.sum(axis=1)
    ga.init(fn)
    best = [ga.fit[0]]
    for _ in range(20):
        before = ga.fit.copy()
        origin = ga.step(fn)
        assert np.all(np.diff(ga.fit) <= 0)  # sorted best-first
        assert ga.fit[0] >= before[0] and ga.fit[-1] >= before[-1]  # (mu + lambda): never worse
        assert origin.min() >= 0 and origin.max() < 128
        best.append(ga.fit[0])
    assert best[-1] > best[0] + 1.0  # it optimises
    assert ga.gen == 20


def test_remap_reevaluates_parents():
    calls = []
    def fn(G):
        calls.append(len(G))
        return G.sum(axis=1)
    ga = OG.NumGA(16, 4, remap=True)
    ga.init(fn)
    ga.step(fn)
    assert calls == [16, 32]


def test_ga_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    from vsrevolution_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        g.build()
    hdr = open(os.path.join(ROOT, "include", "vsr_ga.h")).read()
    declared = set(re.findall(r"\b(vsr_(?:ga|es)_[a-z0-9_]+)\s*\(", hdr))
    lib = C.CDLL(_ffi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in vsr_ga.h but not exported"
    assert declared == set(_ffi.GA_EXPORTED_SYMBOLS)
    # integer rule through the C ABI, no GPU needed
    d = _ffi.GaDesc()
    assert _ffi.lib().vsr_ga_default(d) == 0
    assert (d.xover_prob, d.tournament_rate, d.min_tournament, d.sigma_mut) == (0.75, 0.05, 3, 0.35)  # Starter.java:178
    for n in (1, 10, 41, 61, 100, 4096, 1 << 18):
        d.n_pop = n
        assert _ffi.lib().vsr_ga_tournament_size(d) == OG.tournament_size(n)
    # `ES` constants (Starter.java:179; SimpleES.java:46-47): nOfParents = (int) Math.round(nPop * 0.4), one elite
    for n, npar in ((9, 4), (25, 10), (30, 12), (100, 40), (4096, 1638)):
        assert _ffi.lib().vsr_es_default(d, n) == 0
        assert (d.solver, d.n_pop, d.n_parents, d.n_elites, d.sigma_mut) == (_ffi.SOLVER_ES, n, npar, 1, 0.35)
        assert OG.SimpleES(n, 3).n_parents == npar
    from vsrevolution_b200.solver import SimpleES, es
    eb = es().build({"nPop": "30", "nEval": "300", "remap": "true"})
    assert (eb.n_pop, eb.n_eval, eb.remap, eb.outer.sigma, eb.outer.parents_rate) == (30, 300, True, 0.35, 0.4)
    # the mirror of the builder parses the same parameter names
    from vsrevolution_b200.solver import DoublesStandard, num_ga
    import pytest
    b = num_ga().build({"nPop": "100", "nEval": "1000"})
    assert (b.n_pop, b.n_eval, b.remap) == (100, 1000, False)
    with pytest.raises(_ffi.VsrIllegalArgument):
        DoublesStandard(0.75, 0.05, 3, 0.35).build({"nPop": "100"})
    # diversity=true (StandardWithEnforcedDiversityEvolver) is accepted: with a Gaussian mutation on every offspring its re-draw never fires
    bd = num_ga().build({"nPop": "10", "nEval": "100", "diversity": "true", "remap": "true"})
    assert (bd.n_pop, bd.n_eval, bd.remap) == (10, 100, True)


def test_simple_es_semantics():
    """oracle/ga.py SimpleES = JGEA SimpleEvolutionaryStrategy as builder/solver/SimpleES.java:41-49 configures it: parents = best
    round(0.4 nPop), one elite kept with its fitness, nPop - 1 offspring around the parents' mean, non-overlapping."""
    from oracle import ga as OG
    n, G = 30, 7
    es = OG.SimpleES(n, G, seed=3)
    assert es.n_parents == 12 and es.n_elites == 1
    assert OG.SimpleES(25, G).n_parents == 10 and OG.SimpleES(9, G).n_parents == 4   # Math.round(3.6) = 4, round half up
    sphere = lambda g: -np.sum((g - 0.3) ** 2, axis=1)
    calls = []
    def fit(g):
        calls.append(len(g))
        return sphere(g)
    es.init(fit)
    assert calls == [n] and np.all(np.diff(es.fit) <= 0)
    best0, elite0 = es.fit[0], es.pop[0].copy()
    parents_mean = es.pop[:12].mean(axis=0)
    es.step(fit)
    assert calls == [n, n - 1]                       # the elite is not re-evaluated
    assert np.allclose(es.means, parents_mean, atol=1e-15)
    assert es.fit[0] >= best0 and any(np.array_equal(row, elite0) for row in es.pop)
    # offspring are centred on the means with std sigma
    for _ in range(30):
        es.step(fit)
    assert es.fit[0] > -0.2 and es.fit[0] >= best0       # settles around the optimum (constant sigma: no further)
    m, nxt = es.sample(es.gen + 1)
    z = (nxt[1:] - m) / es.sigma
    assert abs(z.mean()) < 0.2 and 0.8 < z.std() < 1.2
    # remap re-evaluates the elite too
    es2 = OG.SimpleES(n, G, seed=3, remap=True)
    calls.clear()
    es2.init(fit); es2.step(fit)
    assert calls == [n, n]
