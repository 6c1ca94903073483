"""world_size-2 (gloo, CPU) test of the multi-GPU host logic: chain / row shard arithmetic and the end-of-run
moment all-reduce that yields R-hat.  The per-rank moments are synthetic (no GPU here)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from smc_pkg import load_package

load_package()
from simplemcmc_jl_b200 import parallel  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_chains, M, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smc_pkg import load_package
    load_package()
    from simplemcmc_jl_b200 import parallel as par
    rng = np.random.default_rng(123)                      # same stream on every rank: the "global" draws
    draws = rng.standard_normal((n, total_chains, M)) + np.arange(M)
    off, cnt = par.chain_shard(total_chains, rank, world)
    mine = draws[:, off:off + cnt]
    mean = mine.mean(0)
    m2 = ((mine - mean) ** 2).sum(0)
    tot = par.allreduce_sum(par.moment_partials(mean, m2, n))
    out = par.rhat_from_partials(tot, n, M)
    q.put((rank, off, cnt, out["mean"], out["rhat"], out["chains"]))
    dist.destroy_process_group()


def test_chain_shards_cover_all_chains():
    for total, world in [(4096, 8), (10, 4), (7, 2), (3, 8)]:
        spans = [parallel.chain_shard(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
    assert [parallel.row_shard(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]


def test_rhat_allreduce_world_size_2():
    world, total_chains, M, n = 2, 7, 3, 400
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_chains, M, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(123)
    draws = rng.standard_normal((n, total_chains, M)) + np.arange(M)
    cm = draws.mean(0)
    W = draws.var(0, ddof=1).mean(0)
    B = n * cm.var(0, ddof=1)
    rhat = np.sqrt(((n - 1) / n * W + B / n) / W)
    for rank, off, cnt, mean, rh, chains in res:
        assert chains == total_chains
        assert np.allclose(mean, cm.mean(0), rtol=1e-12)
        assert np.allclose(rh, rhat, rtol=1e-10)
