"""cfg3 on a (chain groups x row ranks) grid of GPUs (hybrid row): each row group of ROW_RANKS
consecutive ranks shards the rows of the logistic GLM and advances one block of the chains; the library's NCCL
communicator spans the row group only.  Checks (i) every group's logp/grad == the single-GPU values of its chains to
1e-12, (ii) identical chains inside a row group after HMC, (iii) pooled chain count of the final moment all-reduce, and
times HMC.  Launch (2 x 2): ROW_RANKS=2 python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/hybrid_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api, capi, parallel
from util import LOGISTIC

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lay = parallel.HybridLayout(rank, world, int(os.environ.get("ROW_RANKS", 2)))
N, M, C = int(os.environ.get("ROWS", 2_000_000)), 10, int(os.environ.get("CHAINS", 128))
rng = np.random.default_rng(3)
X = np.empty((M, N)).T
X[:, 0] = 1.0
for j in range(1, M):
    X[:, j] = rng.standard_normal(N)
b0 = rng.standard_normal(M)
Y = (rng.random(N) < 1.0 / (1.0 + np.exp(X @ b0))).astype(np.float64)
r0, r1 = lay.rows(N)
c0, cc = lay.chains(C)
ctx = capi.Context(local)
grp = parallel.init_hybrid_comm(ctx, lay)
low, dev = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=np.asfortranarray(X[r0:r1]), Y=Y[r0:r1]), True, cc, local, "assign",
                      shard="rows" if lay.row_ranks > 1 else None, comm=ctx)
betas = (b0[None, :] + 1e-3 * np.random.default_rng(9).standard_normal((C, M)))[c0:c0 + cc]
lp, g = dev.eval(betas)
ctx1 = capi.Context(local)
low1, dev1 = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=X, Y=Y), True, cc, local, "assign", comm=ctx1)
lp1, g1 = dev1.eval(betas)
dev1.close()
err = torch.tensor([np.max(np.abs(lp - lp1) / np.abs(lp1)), np.max(np.abs(g - g1)) / np.max(np.abs(g1))], device="cuda")
dist.all_reduce(err, op=dist.ReduceOp.MAX)
eps = 0.1 / np.sqrt(N / 1000)
dev.run("hmc", betas, cc, 3, 0, isteps=2, stepsize=eps, seed=1, store_draws=False, want=(), chain_offset=c0)
dist.barrier(); torch.cuda.synchronize()
steps = 20
raw = dev.run("hmc", betas, cc, steps, 0, isteps=2, stepsize=eps, seed=2, store_draws=False, want=(), chain_offset=c0)
ms = torch.tensor([raw["device_ms"]], device="cuda", dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
fs = torch.tensor(raw["final_state"].ravel(), device="cuda")
fmax, fmin = fs.clone(), fs.clone()
dist.all_reduce(fmax, op=dist.ReduceOp.MAX, group=grp); dist.all_reduce(fmin, op=dist.ReduceOp.MIN, group=grp)
same = torch.tensor([1.0 if torch.equal(fmax, fmin) else 0.0], device="cuda")
dist.all_reduce(same, op=dist.ReduceOp.MIN)
tot = parallel.allreduce_sum(parallel.hybrid_moment_partials(lay, raw["mean"], raw["m2"], steps))
pooled = parallel.rhat_from_partials(tot, steps, M)
if rank == 0:
    print(json.dumps(dict(gpus=world, grid="%d chain groups x %d row ranks" % (lay.chain_groups, lay.row_ranks), rows=N, chains=C,
                          logp_relerr_vs_1gpu=float(err[0]), grad_relerr_vs_1gpu=float(err[1]),
                          chains_identical_inside_row_groups=bool(same.item() == 1.0), pooled_chains=pooled["chains"],
                          hmc_ms_per_step=float(ms.item()) / steps,
                          leapfrog_chain_steps_per_s=C * steps * 2 / (float(ms.item()) * 1e-3))))
dist.destroy_process_group()
