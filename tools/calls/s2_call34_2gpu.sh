#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
( time timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_r02.json 2> gpurun_out/bench_2gpu_r02.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_2gpu_r02.json"))
print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, d["e2e"]["value"], d["cpu_baseline"].get("numpy_batched_dgemm_chain_evals_per_s"), d["roofline"].get("shared_pipe_ncu",{}).get("busy_frac"))
x=d["labelled_extras"]; print(sorted(x.keys())); print({k:(v.get("error") if isinstance(v,dict) else None) for k,v in x.items()})
PY
echo "== reference arm under torchrun"; timeout 600 $TR bench.py --impl reference --gpus 2 --steps 4 --warmup 1 2>/dev/null | cut -c1-200
