#!/bin/bash
# session 2, call 41: hierarchical NUTS at the BASELINE size, lockstep only / asynchronous from iteration 0 (is the
# 0.36 switching threshold still right now that an evaluation is 13 launches instead of 21?)
mkdir -p gpurun_out
: > gpurun_out/nuts_baseline_v4d.jsonl
SMC_NUTS_ASYNC=0 timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 2000 2>gpurun_out/nuts_baseline_v4d.err | tee -a gpurun_out/nuts_baseline_v4d.jsonl
SMC_NUTS_ASYNC=1 timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 2000 2>>gpurun_out/nuts_baseline_v4d.err | tee -a gpurun_out/nuts_baseline_v4d.jsonl
tail -3 gpurun_out/nuts_baseline_v4d.err
