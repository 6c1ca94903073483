"""Multi-GPU host logic: one process per GPU (torch.distributed supplies rendezvous, barriers and the small
host-side reductions; the per-evaluation gradient all-reduce of row-sharded models runs inside the library on NCCL).

Chain sharding (BASELINE.json configs 1, 2, 4, 5): chains are independent (no shared state in
/root/reference/src/samplers.jl), rank g runs chains [offset, offset+count) with Philox streams keyed by the global
chain id, data replicated; the only collective is one all-reduce of per-rank chain moments at the end (R-hat).
Row sharding (config 3): rank g holds rows [r0, r1) of X and Y; every rank advances identical chain state.
Hybrid (third row): the ranks form a (chain groups x row ranks) grid; the `row_ranks` consecutive ranks of
a chain group hold the row shards of the data and advance the same chains, the library's NCCL communicator spans only
that group (the per-evaluation all-reduce of (L, G) never leaves it), and one rank per group contributes to the final
moment all-reduce.
"""
import numpy as np


def chain_shard(total_chains, rank, world):
    """contiguous split of the chains; the first `total % world` ranks get one more"""
    base, extra = divmod(int(total_chains), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def row_shard(n_rows, rank, world):
    base, extra = divmod(int(n_rows), int(world))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, start + count


class HybridLayout(object):
    """position of `rank` in the (chain groups x row ranks) grid: consecutive ranks share a chain group, so on one
    NVSwitch box any placement is equivalent and across nodes a row group stays inside a node when row_ranks divides
    the GPUs per node"""

    def __init__(self, rank, world, row_ranks):
        rank, world, row_ranks = int(rank), int(world), int(row_ranks)
        if row_ranks < 1 or world % row_ranks != 0:
            raise ValueError("row_ranks (%d) must divide the number of ranks (%d)" % (row_ranks, world))
        if not 0 <= rank < world:
            raise ValueError("rank %d outside [0, %d)" % (rank, world))
        self.rank, self.world, self.row_ranks = rank, world, row_ranks
        self.chain_groups = world // row_ranks
        self.chain_group, self.row_rank = divmod(rank, row_ranks)

    @property
    def group_ranks(self):
        """global ranks of this rank's row group, in row-rank order"""
        return list(range(self.chain_group * self.row_ranks, (self.chain_group + 1) * self.row_ranks))

    @property
    def group_root(self):
        return self.chain_group * self.row_ranks

    @property
    def counts_moments(self):
        """every rank of a row group holds the same chains: exactly one of them reports their moments"""
        return self.row_rank == 0

    def chains(self, total_chains):
        """(offset, count) of the global chains this rank's group advances"""
        return chain_shard(total_chains, self.chain_group, self.chain_groups)

    def rows(self, n_rows):
        """[start, end) of the data rows this rank holds"""
        return row_shard(n_rows, self.row_rank, self.row_ranks)


def new_row_groups(layout):
    """torch.distributed groups of the grid's row groups.  Collective: every rank creates every group (in the same
    order) and gets its own back."""
    import torch.distributed as dist
    mine = None
    for g in range(layout.chain_groups):
        ranks = list(range(g * layout.row_ranks, (g + 1) * layout.row_ranks))
        grp = dist.new_group(ranks=ranks)
        if g == layout.chain_group:
            mine = grp
    return mine


def moment_partials(chain_mean, chain_m2, n_samples):
    """per-rank sufficient statistics [count, sum mean, sum mean^2, sum var] (3M+1 doubles)"""
    C, M = chain_mean.shape
    var = chain_m2 / max(n_samples - 1, 1)
    return np.concatenate([[float(C)], chain_mean.sum(0), (chain_mean ** 2).sum(0), var.sum(0)])


def rhat_from_partials(total, n_samples, M):
    """Gelman-Rubin potential scale reduction + pooled mean/variance from the all-reduced partials"""
    C = total[0]
    s1, s2, sv = total[1:1 + M], total[1 + M:1 + 2 * M], total[1 + 2 * M:1 + 3 * M]
    mean = s1 / C
    W = sv / C
    B = n_samples * (s2 - C * mean ** 2) / max(C - 1, 1)
    var_plus = (n_samples - 1) / n_samples * W + B / n_samples
    with np.errstate(divide="ignore", invalid="ignore"):
        rhat = np.sqrt(var_plus / W)
    return dict(mean=mean, var=var_plus, rhat=rhat, chains=int(C))


def allreduce_sum(vec, group=None, ctx=None):
    """sum a small float64 vector over ranks: torch.distributed when initialised (gloo on CPU, nccl on GPU),
    otherwise the library's NCCL communicator, otherwise identity (single rank)"""
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    try:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            t = torch.from_numpy(vec.copy())
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, group=group)
            return t.cpu().numpy()
    except ImportError:
        pass
    if ctx is not None and ctx.nranks > 1:
        return ctx.allreduce(vec)
    return vec


def broadcast_unique_id(make_id, rank, group=None, src=0):
    """the 128-byte NCCL unique id of a communicator: the group's rank 0 (`rank` is the rank inside the group, `src`
    that member's global rank) makes it, torch.distributed broadcasts it to the other members"""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src=src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def init_library_comm(ctx, rank, world, group=None, src=0):
    """create the library's NCCL communicator over `group` (default: all ranks); `rank` / `world` are the position in
    and the size of that group, `src` the global rank of its first member"""
    from . import capi
    ctx.comm_init(broadcast_unique_id(capi.unique_id, rank, group=group, src=src), rank, world)


def gather_handles(handle, rank, world, group=None):
    """all ranks' 64-byte handles concatenated in (group) rank order"""
    import torch
    import torch.distributed as dist
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).clone()
    assert mine.numel() == 64
    cuda = dist.get_backend(group) == "nccl"
    if cuda:
        mine = mine.cuda()
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def init_p2p(ctx, rank, world, group=None):
    """peer-memory exchange for a row-sharded model on one node (csrc/smc_p2p.cu): every rank exports the IPC handle of
    its exchange buffer, the handles are gathered in rank order, every rank maps its peers.  `rank` / `world` are the
    position in and the size of `group` (default: all ranks).  Optional: without it the library uses NCCL."""
    handles = gather_handles(ctx.p2p_export(world), rank, world, group=group)
    ctx.p2p_import(handles, world, rank)


def init_hybrid_comm(ctx, layout):
    """library communicator of this rank's row group (hybrid chains x rows sharding); returns the torch group"""
    grp = new_row_groups(layout)
    if layout.row_ranks > 1:
        init_library_comm(ctx, layout.row_rank, layout.row_ranks, group=grp, src=layout.group_root)
    return grp


def hybrid_moment_partials(layout, chain_mean, chain_m2, n_samples):
    """moment_partials for the final all-reduce over ALL ranks of a hybrid grid: the chains of a row group are
    replicated on its members, so only the group's first rank contributes"""
    part = moment_partials(chain_mean, chain_m2, n_samples)
    return part if layout.counts_moments else np.zeros_like(part)
