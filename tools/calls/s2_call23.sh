#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/nuts_baseline_r02.jsonl
for m in sleepstudy hierarchical; do
  timeout 900 python tools/nuts_baseline_run.py $m 16384 2000 2>gpurun_out/nuts_baseline.err | tee -a gpurun_out/nuts_baseline_r02.jsonl
  SMC_NUTS_ASYNC=1 timeout 900 python tools/nuts_baseline_run.py $m 16384 2000 2>>gpurun_out/nuts_baseline.err | tee -a gpurun_out/nuts_baseline_r02.jsonl
done
tail -3 gpurun_out/nuts_baseline.err
