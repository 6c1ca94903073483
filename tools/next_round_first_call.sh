#!/bin/bash
# First GPU call of the next round (about 6 minutes of box time): everything that was written after this round's GPU
# minutes were spent, in the order in which a failure is cheapest.
#   gpurun --timeout 900 -- 'bash tools/next_round_first_call.sh'
mkdir -p gpurun_out
P=tools/ozaki_glm_proto
{
  echo "== prototype v6, v7, v8, v9 (variants 8 to 11): parity, then time at 250k and 1e6 rows"
  timeout 60 $P check 8192 64 3 8
  timeout 60 $P check 8192 64 3 9
  timeout 60 $P check 8192 64 3 10
  timeout 60 $P check 8192 64 3 11
  timeout 60 $P time 250000 4096 37 7
  timeout 60 $P time 250000 4096 37 8
  timeout 60 $P time 250000 4096 37 9
  timeout 60 $P time 250000 4096 37 10
  timeout 60 $P time 250000 4096 37 11
  timeout 120 $P time 1000000 4096 8 7        # 8 row splits: what the library's row_splits_for picks for 128 chain tiles
  echo "== library int8 path: opt-in parity test, then the GLM-related parity tests with the switch set"
  SMC_TEST_I8=1 timeout 900 python -m pytest tests/test_zz_gpu_i8_optin.py -m gpu -x -q 2>&1 | tail -5
  SMC_GLM_I8=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_samplers.py -m gpu -q 2>&1 | tail -8
  echo "== library int8 path at configs[1]: rate, event-timed launch, parity of four chains (7 and 8 beta slices)"
  SMC_GLM_I8=1 timeout 300 python tools/i8_bench_extra.py 10 3
  SMC_GLM_I8=1 SMC_GLM_I8_BETA8=1 timeout 300 python tools/i8_bench_extra.py 10 3
  SMC_GLM_I8=1 SMC_GLM_I8_BETA8=1 timeout 60 python tools/i8_quick.py
  echo "== launch list of one evaluation with the switch set"
  SMC_GLM_I8=1 SMC_NO_GRAPHS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches_i8.csv python tools/i8_quick.py > gpurun_out/ncu_i8.log 2>&1; tail -2 gpurun_out/ncu_i8.log
} > gpurun_out/next_round_first_call.log 2>&1
tail -60 gpurun_out/next_round_first_call.log

# Second call, two GPUs (charged twice): NCCL against the peer-memory exchange of csrc/smc_p2p.cu on a latency-bound shape, and
# the hybrid grid on four:
#   gpurun --gpus 2 --timeout 600 -- 'for p in 0 1; do P2P=$p ROWS=2000000 CHAINS=2 python -m torch.distributed.run --nproc-per-node 2 \
#       --master-addr 127.0.0.1 tools/rowshard_check.py; done > gpurun_out/p2p_2gpu.log 2>&1; tail -4 gpurun_out/p2p_2gpu.log'
#   gpurun --gpus 4 --timeout 600 -- 'ROW_RANKS=2 python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 \
#       tools/hybrid_check.py > gpurun_out/hybrid_4gpu.log 2>&1; tail -2 gpurun_out/hybrid_4gpu.log'
