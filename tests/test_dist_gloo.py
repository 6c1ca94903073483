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


# ---- hybrid (chain groups x row ranks) grid, SURVEY.md 8e third row: world_size 4 = 2 x 2 on gloo ------------------

def _glm_partial(X, Y, betas):
    """row-partial (L, G) of Y - X*beta ~ Normal(0, 1) for a block of chains: what the row-sharded GLM statement sums
    over the row group"""
    resid = Y[:, None] - X @ betas.T                       # rows x chains
    L = np.sum(-0.5 * resid ** 2 - 0.5 * np.log(2 * np.pi), axis=0)
    G = (X.T @ resid).T                                    # chains x M
    return np.concatenate([L[:, None], G], axis=1)


def _hybrid_problem():
    rng = np.random.default_rng(77)
    N, M, C, n = 101, 3, 6, 50
    X = rng.standard_normal((N, M))
    Y = X @ rng.standard_normal(M) + rng.standard_normal(N)
    betas = rng.standard_normal((C, M))
    draws = rng.standard_normal((n, C, M)) + np.arange(M)
    return N, M, C, n, X, Y, betas, draws


def _hybrid_worker(rank, world, row_ranks, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smc_pkg import load_package
    load_package()
    from simplemcmc_jl_b200 import parallel as par
    lay = par.HybridLayout(rank, world, row_ranks)
    grp = par.new_row_groups(lay)
    # the id exchange init_library_comm performs, with a stand-in for ncclGetUniqueId (no GPU here)
    uid = par.broadcast_unique_id(lambda: bytes([17 + lay.chain_group]) * 128, lay.row_rank, group=grp, src=lay.group_root)
    N, M, C, n, X, Y, betas, draws = _hybrid_problem()
    c0, cc = lay.chains(C)
    r0, r1 = lay.rows(N)
    # per-evaluation exchange: (L, G) of my rows for my group's chains, summed inside the row group only
    LG = par.allreduce_sum(_glm_partial(X[r0:r1], Y[r0:r1], betas[c0:c0 + cc]).ravel(), group=grp).reshape(cc, M + 1)
    # end of run: chain moments over ALL ranks, one contributor per row group
    mine = draws[:, c0:c0 + cc]
    mean = mine.mean(0)
    m2 = ((mine - mean) ** 2).sum(0)
    tot = par.allreduce_sum(par.hybrid_moment_partials(lay, mean, m2, n))
    out = par.rhat_from_partials(tot, n, M)
    q.put((rank, lay.chain_group, lay.row_rank, uid[:1], (c0, cc), (r0, r1), LG, out["mean"], out["rhat"], out["chains"]))
    dist.barrier()
    dist.destroy_process_group()


def test_hybrid_layout_arithmetic():
    import pytest
    for world, rr in [(8, 2), (8, 4), (4, 2), (8, 1), (8, 8), (6, 3)]:
        lays = [parallel.HybridLayout(r, world, rr) for r in range(world)]
        assert sorted(sum((l.group_ranks for l in lays if l.row_rank == 0), [])) == list(range(world))
        assert sum(l.counts_moments for l in lays) == world // rr
        for l in lays:
            assert l.rank == l.chain_group * rr + l.row_rank and l.group_root == l.group_ranks[0]
        # chains: every chain in exactly one group; rows: every row on exactly one member of each group
        C, N = 37, 1001
        owners = np.zeros(C, int)
        for l in lays:
            if l.row_rank == 0:
                o, c = l.chains(C)
                owners[o:o + c] += 1
        assert np.all(owners == 1)
        cover = np.zeros((world // rr, N), int)
        for l in lays:
            a, b = l.rows(N)
            cover[l.chain_group, a:b] += 1
        assert np.all(cover == 1)
    with pytest.raises(ValueError):
        parallel.HybridLayout(0, 8, 3)
    with pytest.raises(ValueError):
        parallel.HybridLayout(8, 8, 2)


def test_hybrid_grid_world_size_4():
    world, row_ranks = 4, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hybrid_worker, args=(r, world, row_ranks, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, M, C, n, X, Y, betas, draws = _hybrid_problem()
    full = _glm_partial(X, Y, betas)
    cm = draws.mean(0)
    W = draws.var(0, ddof=1).mean(0)
    B = n * cm.var(0, ddof=1)
    rhat = np.sqrt(((n - 1) / n * W + B / n) / W)
    by_group = {}
    for rank, cg, rr, uid0, (c0, cc), (r0, r1), LG, mean, rh, chains in res:
        assert uid0 == bytes([17 + cg])                    # each row group got its own leader's id
        assert np.allclose(LG, full[c0:c0 + cc], rtol=1e-12, atol=1e-12)
        by_group.setdefault(cg, []).append(LG)
        assert chains == C                                 # replicated chains counted once
        assert np.allclose(mean, cm.mean(0), rtol=1e-12)
        assert np.allclose(rh, rhat, rtol=1e-10)
    for cg, lgs in by_group.items():
        assert len(lgs) == row_ranks and np.array_equal(lgs[0], lgs[1])     # identical on the members of a row group


# ---- peer-memory handle exchange (host side of csrc/smc_p2p.cu): world_size 2 on gloo, with a stand-in context -------------

class _FakeCtx(object):
    def __init__(self, rank):
        self.rank, self.imported = rank, None

    def p2p_export(self, nranks):
        return bytes([100 + self.rank]) * 64

    def p2p_import(self, handles, nranks, rank):
        self.imported = (handles, nranks, rank)


def _p2p_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smc_pkg import load_package
    load_package()
    from simplemcmc_jl_b200 import parallel as par
    ctx = _FakeCtx(rank)
    par.init_p2p(ctx, rank, world)
    q.put((rank, ctx.imported))
    dist.barrier()
    dist.destroy_process_group()


def test_p2p_handles_are_gathered_in_rank_order():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, (handles, nranks, r) in res:
        assert (nranks, r) == (world, rank)
        assert handles == bytes([100]) * 64 + bytes([101]) * 64
