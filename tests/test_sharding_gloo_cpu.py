"""N>1 host-side path on CPU: world_size-2 gloo run of the sharding + gather plumbing bench.py uses
(robots split by contiguous index range, no data-path collective; fitness gathered afterwards).
The per-rank evaluator is the ORACLE here (tests may use it); on the GPU box it is the CUDA evaluator."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import model as O
    from vsrevolution_b200.locomotion import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = O.config_model(1, final_t=1.5)
    g = np.random.default_rng(0).uniform(-1, 1, (n, m.genome_len()))  # same population on every rank
    lo, hi = shard_range(n, rank, world)
    mine = m.eval_batch(g[lo:hi]) if hi > lo else np.zeros(0)
    per = (n + world - 1) // world
    buf = torch.full((per,), float("nan"), dtype=torch.float64)
    buf[: hi - lo] = torch.from_numpy(mine)
    gathered = [torch.empty(per, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, buf)
    full = torch.cat([gathered[r][: shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)]).numpy()
    if rank == 0:
        out_q.put(full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8])
def test_two_rank_sharded_evaluation_equals_single_rank(n):
    import torch.multiprocessing as mp
    from oracle import model as O
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = O.config_model(1, final_t=1.5)
    g = np.random.default_rng(0).uniform(-1, 1, (n, m.genome_len()))
    assert np.array_equal(full, m.eval_batch(g))


def _ga_worker(rank, world, port, n_pop, out_q, solver="ga"):
    """The N>1 generation loop (vsrevolution_b200/solver.py): replicated GA state, sharded evaluation, fitness all-gather.
    On CPU the GA is oracle/ga.py and the per-rank evaluator a cheap deterministic function; on the GPU box the same exchange
    runs between vsr_ga_step_begin and vsr_ga_step_end over NCCL."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import ga as OG
    from vsrevolution_b200.locomotion import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fn = lambda G: np.sin(G).sum(axis=1) - 0.1 * (G * G).sum(axis=1)

    def sharded(G):
        n = len(G)
        per = (n + world - 1) // world
        lo, hi = shard_range(n, rank, world)
        buf = torch.zeros(per, dtype=torch.float64)
        buf[: hi - lo] = torch.from_numpy(fn(G[lo:hi]))
        whole = torch.empty(per * world, dtype=torch.float64)
        dist.all_gather_into_tensor(whole, buf)
        return torch.cat([whole[r * per: r * per + (shard_range(n, r, world)[1] - shard_range(n, r, world)[0])] for r in range(world)]).numpy()

    ga = OG.NumGA(n_pop, 12, seed=9) if solver == "ga" else OG.SimpleES(n_pop, 12, seed=9)
    ga.init(sharded)
    for _ in range(4):
        ga.step(sharded)  # ES evaluates n_pop - 1 offspring: an odd count split over the ranks
    if rank == 0:
        out_q.put((ga.pop.copy(), ga.fit.copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pop,solver", [(25, "ga"), (32, "ga"), (25, "es"), (32, "es")])
def test_two_rank_generation_loop_equals_single_rank(n_pop, solver):
    import torch.multiprocessing as mp
    from oracle import ga as OG
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ga_worker, args=(r, 2, port, n_pop, q, solver)) for r in range(2)]
    for p in procs:
        p.start()
    pop, fit = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fn = lambda G: np.sin(G).sum(axis=1) - 0.1 * (G * G).sum(axis=1)
    ga = OG.NumGA(n_pop, 12, seed=9) if solver == "ga" else OG.SimpleES(n_pop, 12, seed=9)
    ga.init(fn)
    for _ in range(4):
        ga.step(fn)
    assert np.array_equal(pop, ga.pop) and np.array_equal(fit, ga.fit)
