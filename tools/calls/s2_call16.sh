#!/bin/bash
# session 2, call 16: simpleNUTS with asynchronous iterations: sampler tests, whole gpu suite, cfg4 sweep (default and SMC_NUTS_LOCKSTEP=1)
mkdir -p gpurun_out
echo "== sampler tests"; timeout 1200 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -12
echo "== cfg4 sweep (asynchronous iterations)"; timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_async.json 2>gpurun_out/sweep_cfg4_async.err; grep -E "chains|leapfrogs_per_s|active_lane|evals" gpurun_out/sweep_cfg4_async.json | paste - - - - | tail -4; tail -2 gpurun_out/sweep_cfg4_async.err
echo "== cfg4 sweep (lockstep)"; SMC_NUTS_LOCKSTEP=1 timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_lockstep.json 2>/dev/null; grep -E "chains|leapfrogs_per_s|active_lane|evals" gpurun_out/sweep_cfg4_lockstep.json | paste - - - - | tail -4
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
