#!/bin/bash
# session 2, call 40: products with one left operand paired into one launch: smoke, whole gpu suite,
# simpleNUTS at the BASELINE size in the default mode
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) 2>&1
: > gpurun_out/nuts_baseline_v4c.jsonl
for m in sleepstudy hierarchical; do
  timeout 600 python tools/nuts_baseline_run.py $m 16384 2000 2>gpurun_out/nuts_baseline_v4c.err | tee -a gpurun_out/nuts_baseline_v4c.jsonl
done
tail -3 gpurun_out/nuts_baseline_v4c.err
