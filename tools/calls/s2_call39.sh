#!/bin/bash
# session 2, call 39: tape version 4 complete (accumulator zero between evaluations), one-block slot packing, unsplit products stored directly: smoke, whole gpu suite,
# simpleNUTS at the BASELINE size in the default mode
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) 2>&1
: > gpurun_out/nuts_baseline_v4b.jsonl
for m in sleepstudy hierarchical; do
  timeout 600 python tools/nuts_baseline_run.py $m 16384 2000 2>gpurun_out/nuts_baseline_v4b.err | tee -a gpurun_out/nuts_baseline_v4b.jsonl
done
tail -3 gpurun_out/nuts_baseline_v4b.err
