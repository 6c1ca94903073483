#!/bin/bash
mkdir -p gpurun_out
echo "== sampler tests"; timeout 1200 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -6
echo "== cfg4 sweep (default = probe, then decide)"; timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_auto.json 2>gpurun_out/sweep_cfg4_auto.err; grep -E "chains|leapfrogs_per_s|active_lane|evals" gpurun_out/sweep_cfg4_auto.json | paste - - - - | tail -4; tail -2 gpurun_out/sweep_cfg4_auto.err
echo "== BASELINE size, hierarchical, default policy (600 iterations)"; timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 600 2>&1 | tail -1
