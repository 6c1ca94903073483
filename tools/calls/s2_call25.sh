#!/bin/bash
mkdir -p gpurun_out
echo "== sampler tests"; timeout 1200 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -4
echo "== cfg4 sweep (default policy)"; timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_auto.json 2>gpurun_out/sweep_cfg4_auto.err; grep -E "chains|leapfrogs_per_s|active_lane|evals" gpurun_out/sweep_cfg4_auto.json | paste - - - - | tail -4; tail -2 gpurun_out/sweep_cfg4_auto.err
echo "== hierarchical, default policy, 16384 x 600"; timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 600 2>&1 | tail -1
echo "== sleepstudy, default policy, 16384 x 300"; timeout 600 python tools/nuts_baseline_run.py sleepstudy 16384 300 2>&1 | tail -1
