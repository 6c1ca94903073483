#!/bin/bash
# session 2, call 42: the final tree -- smoke, whole gpu suite, bench line (with the bounded simpleNUTS extra)
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err ) 2>&1 | tail -3; tail -3 gpurun_out/bench_1gpu_r02.err; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_r02.json')); print(d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['traffic']); print(json.dumps(d['labelled_extras'].get('nuts_cfg4_16k_chains'))[:900]); print(d['labelled_extras']['cfg1_2p20_chains'].get('value'))"
