#!/bin/bash
# session 2, call 44: k_slot_pack with thread-contiguous ranges and a block scan: sampler tests, simpleNUTS at the BASELINE size
mkdir -p gpurun_out
echo "== sampler tests"; timeout 600 python -m pytest tests/test_gpu_samplers.py -m gpu -q 2>&1 | tail -4
: > gpurun_out/nuts_baseline_v4e.jsonl
for m in sleepstudy hierarchical; do
  timeout 300 python tools/nuts_baseline_run.py $m 16384 2000 2>gpurun_out/nuts_baseline_v4e.err | tee -a gpurun_out/nuts_baseline_v4e.jsonl | cut -c1-330
done
tail -2 gpurun_out/nuts_baseline_v4e.err
