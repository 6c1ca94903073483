#!/bin/bash
# session 2, call 45 (2 GPUs, the last GPU-seconds of the round): rows sharded over two ranks with the tape-version-4 library,
# peer-memory exchange: logp / gradient against the whole data on one GPU, identical chains on both ranks after HMC
mkdir -p gpurun_out
ROWS=400000 CHAINS=64 P2P=${P2P:-1} timeout 65 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 tools/rowshard_check.py 2>gpurun_out/rowshard_v4.err | tee gpurun_out/rowshard_v4_p2p${P2P:-1}.json
tail -2 gpurun_out/rowshard_v4.err
