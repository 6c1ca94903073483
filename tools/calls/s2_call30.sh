#!/bin/bash
mkdir -p gpurun_out
echo "== sampler tests"; timeout 1200 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -3
echo "== cfg4 sweep"; timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_r02b.json 2>/dev/null; grep -E "chains|leapfrogs_per_s|active_lane" gpurun_out/sweep_cfg4_r02b.json | paste - - - | tail -4
: > gpurun_out/nuts_baseline_r02b.jsonl
for m in sleepstudy hierarchical; do timeout 900 python tools/nuts_baseline_run.py $m 16384 2000 2>/dev/null | tail -1 | tee -a gpurun_out/nuts_baseline_r02b.jsonl | cut -c1-400; done
