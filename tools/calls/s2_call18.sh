#!/bin/bash
# session 2, call 18: the committed tree once more -- smoke, whole gpu suite, bench line
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err ) 2>&1 | tail -3; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_r02.json')); print(d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['traffic'])"
