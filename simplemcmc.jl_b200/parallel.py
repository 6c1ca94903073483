"""Multi-GPU host logic: one process per GPU (torch.distributed supplies rendezvous, barriers and the small
host-side reductions; the per-evaluation gradient all-reduce of row-sharded models runs inside the library on NCCL).

Chain sharding (BASELINE.json configs 1, 2, 4, 5): chains are independent (no shared state in
/root/reference/src/samplers.jl), rank g runs chains [offset, offset+count) with Philox streams keyed by the global
chain id, data replicated; the only collective is one all-reduce of per-rank chain moments at the end (R-hat).
Row sharding (config 3): rank g holds rows [r0, r1) of X and Y; every rank advances identical chain state.
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


def init_library_comm(ctx, rank, world, group=None):
    """create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed broadcasts it"""
    import torch
    import torch.distributed as dist
    from . import capi
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(capi.unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src=0, group=group)
    ctx.comm_init(bytes(buf.cpu().numpy().tobytes()), rank, world)
