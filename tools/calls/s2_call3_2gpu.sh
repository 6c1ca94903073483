#!/bin/bash
# session 2, call 3 (two GPUs): the peer-memory exchange against NCCL on the latency-bound shape, then the bench line with its N > 1 extras
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for p in 1 0; do
  echo "== rowshard_check P2P=$p, 2 chains"; P2P=$p ROWS=10000000 CHAINS=2 timeout 400 $TR tools/rowshard_check.py 2>&1 | tail -3
done
echo "== bench --gpus 2"; ( time timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_s2c3.json 2> gpurun_out/bench_2gpu_s2c3.err ) 2>&1 | tail -4
tail -c 6000 gpurun_out/bench_2gpu_s2c3.json; tail -5 gpurun_out/bench_2gpu_s2c3.err
