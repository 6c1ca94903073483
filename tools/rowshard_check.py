"""cfg3 on G GPUs (torchrun): rows of the logistic GLM sharded over ranks, per-evaluation NCCL all-reduce of (L, G)
inside the library.  Checks (i) sharded logp/grad == single-GPU logp/grad to 1e-12, (ii) identical chains on every rank,
and times HMC.  Launch: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/rowshard_check.py"""
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
N, M, C = int(os.environ.get("ROWS", 10_000_000)), 10, int(os.environ.get("CHAINS", 64))
rng = np.random.default_rng(3)
X = np.empty((M, N)).T
X[:, 0] = 1.0
for j in range(1, M):
    X[:, j] = rng.standard_normal(N)
b0 = rng.standard_normal(M)
Y = (rng.random(N) < 1.0 / (1.0 + np.exp(X @ b0))).astype(np.float64)
r0, r1 = parallel.row_shard(N, rank, world)
ctx = capi.Context(local)
parallel.init_library_comm(ctx, rank, world)
if os.environ.get("P2P") == "1":      # peer-memory exchange instead of the NCCL all-reduce (csrc/smc_p2p.cu)
    parallel.init_p2p(ctx, rank, world)
low, dev = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=np.asfortranarray(X[r0:r1]), Y=Y[r0:r1]), True, C, local, "assign",
                      shard="rows", comm=ctx)
betas = b0[None, :] + 1e-3 * np.random.default_rng(9).standard_normal((C, M))
lp, g = dev.eval(betas)
res = {}
if rank == 0:
    ctx1 = capi.Context(local)
    low1, dev1 = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=X, Y=Y), True, C, local, "assign", comm=ctx1)
    lp1, g1 = dev1.eval(betas)
    res["logp_relerr_vs_1gpu"] = float(np.max(np.abs(lp - lp1) / np.abs(lp1)))
    res["grad_relerr_vs_1gpu"] = float(np.max(np.abs(g - g1)) / np.max(np.abs(g1)))
    dev1.close()
t = torch.tensor(np.concatenate([lp, g.ravel()]), device="cuda")
tmax, tmin = t.clone(), t.clone()
dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
res["identical_on_all_ranks"] = bool(torch.equal(tmax, tmin))
eps = 0.1 / np.sqrt(N / 1000)
dev.run("hmc", betas, C, 3, 0, isteps=2, stepsize=eps, seed=1, store_draws=False, want=())
dist.barrier(); torch.cuda.synchronize()
raw = dev.run("hmc", betas, C, 10, 0, isteps=2, stepsize=eps, seed=2, store_draws=False, want=())
ms = torch.tensor([raw["device_ms"]], device="cuda", dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
fs = torch.tensor(raw["final_state"].ravel(), device="cuda")
fmax, fmin = fs.clone(), fs.clone()
dist.all_reduce(fmax, op=dist.ReduceOp.MAX); dist.all_reduce(fmin, op=dist.ReduceOp.MIN)
res.update(exchange="p2p" if os.environ.get("P2P") == "1" else "nccl", gpus=world, rows=N, chains=C, hmc_ms_per_step=float(ms.item()) / 10, accept=float(raw["accept_rate"].mean()),
           leapfrog_chain_steps_per_s=C * 20 / (float(ms.item()) * 1e-3), chains_identical_after_hmc=bool(torch.equal(fmax, fmin)))
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
