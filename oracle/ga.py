"""CPU ORACLE of the generation loop (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

numpy restatement of the solver the reference names ``numGA``:
``new DoublesStandard(0.75, 0.05, 3, 0.35)`` (/root/reference/.../locomotion/Starter.java:178), built in
/root/reference/.../builder/solver/DoublesStandard.java:41-71 as a JGEA ``StandardEvolver``:

* factory ``FixedLengthListFactory(G, UniformDoubleFactory(-1, 1))``                             (:52-53)
* operators ``GaussianMutation(sigma)`` with weight 1-pX, ``UniformCrossover.andThen(GaussianMutation)`` with weight pX  (:54-57)
* parent selector ``Tournament(max(minNTournament, ceil(nPop * tournamentRate)))``               (:65)
* unsurvival selector ``Last()``, ``nOffspring = nPop``, ``overlapping = true``                  (:66-69)

PARITY UNPINNED against JGEA 2.0.1 (un-vendored, pom.xml:23-27): JGEA draws everything from one sequential
``java.util.Random``; the device loop (include/vsr_ga.h) uses counter-based Philox4x32-10 draws instead, so the parity
claim is *semantics*, and this file is the index-for-index statement of those semantics.  Philox itself is pinned by the
Random123 known-answer vectors (tests/test_ga_oracle_cpu.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np

S_INIT, S_OP, S_PARENT1, S_PARENT2, S_XOVER, S_MUT = range(6)
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32(key, ctr):
    """Philox4x32-10.  key: (k0, k1) ints; ctr: 4 broadcastable integer arrays -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & _U32 for x in np.broadcast_arrays(*ctr)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = np.uint64(_M0) * c[0]
        p1 = np.uint64(_M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0)) & _U32, p1 & _U32,
             ((p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1)) & _U32, p0 & _U32]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def u53(hi, lo):
    """[0, 1) from 53 random bits, the layout csrc/vsr_ga.cu uses."""
    k = ((hi.astype(np.uint64) >> np.uint64(5)) << np.uint64(26)) | (lo.astype(np.uint64) >> np.uint64(6))
    return k.astype(np.float64) * (1.0 / 9007199254740992.0)


def tournament_size(n_pop: int, rate: float = 0.05, min_n: int = 3) -> int:
    return max(min_n, int(math.ceil(n_pop * rate)))  # DoublesStandard.java:65


def order(fitness: np.ndarray) -> np.ndarray:
    """Indices best-first: higher fitness first (comparator Starter.java:434), NaN as -inf, ties -> lower index."""
    f = np.where(np.isnan(fitness), -np.inf, fitness)
    return np.lexsort((np.arange(len(f)), -f))


class NumGA:
    def __init__(self, n_pop: int, genome_len: int, seed: int = 0, xover_prob: float = 0.75, tournament_rate: float = 0.05,
                 min_tournament: int = 3, sigma_mut: float = 0.35, init_range=(-1.0, 1.0), remap: bool = False):
        self.n, self.G = n_pop, genome_len
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.px, self.sigma, self.rng = xover_prob, sigma_mut, init_range
        self.nt = tournament_size(n_pop, tournament_rate, min_tournament)
        self.remap = remap
        self.gen = -1
        self.pop = self.fit = None

    # -- generation 0 ------------------------------------------------------------------------------------------------------
    def initial_genomes(self) -> np.ndarray:
        i = np.arange(self.n)[:, None]
        g = np.arange(self.G)[None, :]
        r = philox4x32(self.key, (0, i, S_INIT, g))
        return self.rng[0] + (self.rng[1] - self.rng[0]) * u53(r[0], r[1])

    def set_population(self, genomes: np.ndarray, fitness: np.ndarray) -> np.ndarray:
        o = order(fitness)
        self.pop, self.fit = np.array(genomes[o]), np.array(fitness[o])
        self.gen = max(self.gen, 0)
        return o

    # -- one generation ----------------------------------------------------------------------------------------------------
    def select(self, gen: int) -> np.ndarray:
        """parents [n][2]; column 1 = -1 for mutation-only offspring.  Population is sorted => tournament winner = min index."""
        i = np.arange(self.n)
        op = philox4x32(self.key, (gen, i, S_OP, 0))[0]
        xover = op.astype(np.float64) * (1.0 / 4294967296.0) < self.px
        out = np.empty((self.n, 2), dtype=np.int32)
        nblk = (self.nt + 3) // 4
        for slot, stream in ((0, S_PARENT1), (1, S_PARENT2)):
            r = philox4x32(self.key, (gen, i[:, None], stream, np.arange(nblk)[None, :]))
            draws = np.stack(r, axis=-1).reshape(self.n, nblk * 4)[:, : self.nt]
            idx = (draws.astype(np.uint64) * np.uint64(self.n)) >> np.uint64(32)
            out[:, slot] = idx.min(axis=1).astype(np.int32)
        out[~xover, 1] = -1
        return out

    def vary(self, gen: int, parents: np.ndarray) -> np.ndarray:
        i = np.arange(self.n)[:, None]
        g = np.arange(self.G)[None, :]
        x = philox4x32(self.key, (gen, i, S_XOVER, g >> 7))
        words = np.stack(x, axis=-1)  # [n][G][4]
        w = np.take_along_axis(words, ((g >> 5) & 3)[..., None].repeat(self.n, axis=0), axis=-1)[..., 0]
        coin = ((w >> (g & 31).astype(np.uint32)) & np.uint32(1)).astype(bool)
        p1, p2 = parents[:, 0:1], parents[:, 1:2]
        src = np.where((p2 >= 0) & coin, p2, p1)
        base = self.pop[src, g]
        r = philox4x32(self.key, (gen, i, S_MUT, g))
        u1 = u53(r[0], r[1]) + 1.0 / 9007199254740992.0
        u2 = u53(r[2], r[3])
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586 * u2)
        return base + self.sigma * z

    def merge(self, off: np.ndarray, off_fit: np.ndarray, pop_fit: np.ndarray | None = None) -> np.ndarray:
        """(mu + lambda) truncation; returns origin rows (0..n-1 = old population, n..2n-1 = offspring)."""
        allg = np.concatenate([self.pop, off])
        allf = np.concatenate([self.fit if pop_fit is None else pop_fit, off_fit])
        o = order(allf)[: self.n]
        self.pop, self.fit = allg[o], allf[o]
        self.gen += 1
        return o

    def step(self, fitness_fn):
        parents = self.select(self.gen + 1)
        off = self.vary(self.gen + 1, parents)
        if self.remap:
            f = fitness_fn(np.concatenate([self.pop, off]))
            return self.merge(off, f[self.n:], f[: self.n])
        return self.merge(off, fitness_fn(off))

    def init(self, fitness_fn):
        g = self.initial_genomes()
        return self.set_population(g, fitness_fn(g))


class SimpleES:
    """The solver the reference names ``ES``: ``new SimpleES(0.35, 0.4)`` (locomotion/Starter.java:179), built in
    builder/solver/SimpleES.java:41-49 as a JGEA ``SimpleEvolutionaryStrategy(mapper, UniformDoubleFactory(-1, 1) genotypes, nPop,
    stop, nOfParents = round(nPop * parentsRate), nOfElites = 1, sigma, remap)``.  One generation [U: JGEA 2.0.1 is un-vendored]:
    sort; means = gene-wise mean of the best nOfParents; next population = the nOfElites best (fitness kept, or re-evaluated with
    remap) + (nPop - nOfElites) offspring ``means + N(0, 1) * sigma``.  Draws: Philox stream S_MUT at (generation, offspring, gene)."""

    def __init__(self, n_pop: int, genome_len: int, seed: int = 0, sigma: float = 0.35, parents_rate: float = 0.4,
                 n_elites: int = 1, init_range=(-1.0, 1.0), remap: bool = False):
        self.n, self.G = n_pop, genome_len
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.sigma, self.rng, self.remap = sigma, init_range, remap
        self.n_parents = int(math.floor(n_pop * parents_rate + 0.5))  # (int) Math.round(...)
        self.n_elites = n_elites
        self.gen = -1
        self.pop = self.fit = self.means = None

    initial_genomes = NumGA.initial_genomes
    set_population = NumGA.set_population

    def init(self, fitness_fn):
        g = self.initial_genomes()
        return self.set_population(g, fitness_fn(g))

    def sample(self, gen: int):
        """(means, next-generation genomes [n][G]: elites first, then the offspring in draw order)"""
        s = np.zeros(self.G)
        for i in range(self.n_parents):  # rank order, like the device's es_means_kernel
            s = s + self.pop[i]
        means = s / float(self.n_parents)
        e = self.n_elites
        i = np.arange(self.n - e)[:, None]
        g = np.arange(self.G)[None, :]
        r = philox4x32(self.key, (gen, i, S_MUT, g))
        u1 = u53(r[0], r[1]) + 1.0 / 9007199254740992.0
        u2 = u53(r[2], r[3])
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(6.283185307179586 * u2)
        return means, np.concatenate([self.pop[:e], means[None, :] + self.sigma * z])

    def step(self, fitness_fn):
        self.means, nxt = self.sample(self.gen + 1)
        e = self.n_elites
        if self.remap:
            f = fitness_fn(nxt)
        else:
            f = np.concatenate([self.fit[:e], fitness_fn(nxt[e:])])
        o = order(f)
        self.pop, self.fit = nxt[o], f[o]
        self.gen += 1
        return o
