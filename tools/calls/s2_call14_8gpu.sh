#!/bin/bash
# session 2, call 14 (eight GPUs): the bench line with its N > 1 extras on the whole box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
( time timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_r02.json 2> gpurun_out/bench_8gpu_r02.err ) 2>&1 | tail -4
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_8gpu_r02.json"))
print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, d["e2e"]["value"])
x=d["labelled_extras"]
print("cfg1", x["cfg1_2p20_chains"].get("value"), x["cfg1_2p20_chains"].get("error"))
print("strong", x["strong_scaling_cfg2"])
rs=x["row_sharded_cfg3"]
for k,v in rs.items():
    if isinstance(v,dict): print(k, {a:v[a] for a in ("hmc_ms_per_step","leapfrog_us","speedup_vs_1gpu","logp_relerr_vs_1gpu","grad_relerr_vs_1gpu","chains_identical_on_all_ranks_after_hmc") if a in v})
    elif k=="error": print("ERROR", v)
PY
tail -3 gpurun_out/bench_8gpu_r02.err
