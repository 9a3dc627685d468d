"""N > 1 host logic on CPU (gloo, world_size 2): strided sharding of a minibatch of windows
plus ONE sum all-reduce of the packed statistics reproduces the single-process statistics.
The per-rank statistics here come from the oracle (checker only; the GPU path is covered by
the -m gpu tests), what is under test is pysvihmm_b200.sharding."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _packed(O, p, starts, T):
    K, D = p["var_tran"].shape[0], p["obs"].shape[1]
    if len(starts) == 0:
        return np.zeros(K * K + K + K * D + K * D * D + K + 4)
    r = O.svi_minibatch_step(p["obs"], p["mask"], starts, T, p["var_tran"], p["emit"], p["prior_tran"],
                             p["prior_emit"], 0.5, T // 2)
    n = np.array([e[1] for e in r["emit_inter"]])
    sx = np.concatenate([e[0] for e in r["emit_inter"]])
    sxx = np.concatenate([e[2].ravel() for e in r["emit_inter"]])
    tail = np.array([r["logZ"].sum(), r["lb"], len(starts), 0.])
    return np.concatenate([r["A_inter"].ravel(), n, sx, sxx, r["var_x"][:, 0].sum(0), tail])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import svihmm_oracle as O
    from pysvihmm_b200.sharding import allreduce_stats, dist_or_none, shard_starts
    from tests.helpers import make_random_problem
    p = make_random_problem(seed=2, K=4, D=2, T_full=300, miss=0.1)
    starts = np.random.RandomState(0).randint(0, 300 - 21 + 1, 7)       # odd: ranks get 4 and 3
    mine = shard_starts(starts, rank, world)
    assert dist_or_none() is not None
    stats = torch.from_numpy(_packed(O, p, mine, 21))
    allreduce_stats(stats)
    full = _packed(O, p, starts, 21)
    ok = np.allclose(stats.numpy(), full, rtol=1e-12, atol=1e-12) and stats.numpy()[-2] == 7
    out[rank] = bool(ok) and len(mine) == (4 if rank == 0 else 3)
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(0) is True and out.get(1) is True


def _worker_bcast(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pysvihmm_b200.sharding import broadcast_minibatch, shard_starts
    # each rank drew its OWN minibatch and window length (seed=None case): rank 0's must win
    mine = np.random.RandomState(100 + rank).randint(0, 1000, 5 + rank)
    starts, T = broadcast_minibatch(mine, 21 + 2 * rank, dist, torch.device("cpu"))
    ref = np.random.RandomState(100).randint(0, 1000, 5)
    ok = T == 21 and np.array_equal(starts, ref)
    # a minibatch smaller than the world: the last rank's shard is empty, the others are not
    one, _ = broadcast_minibatch(ref[:1], 21, dist, torch.device("cpu"))
    sh = shard_starts(one, rank, world)
    out[rank] = bool(ok) and len(sh) == (1 if rank == 0 else 0)
    dist.destroy_process_group()


def test_broadcast_minibatch_and_empty_shard():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker_bcast, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(0) is True and out.get(1) is True


def test_shard_starts_partitions():
    from pysvihmm_b200.sharding import shard_starts
    s = np.arange(13)
    parts = [shard_starts(s, r, 4) for r in range(4)]
    assert sorted(np.concatenate(parts)) == list(s)
    assert max(map(len, parts)) - min(map(len, parts)) <= 1
