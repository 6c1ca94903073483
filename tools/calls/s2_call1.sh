#!/bin/bash
# session 2, call 1: the whole gpu suite with the int8 path as default, the bench line, the cfg5 sweep with the specialised AFFNORM kernels
mkdir -p gpurun_out
echo "== gpu suite"; ( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_s2c1.json 2> gpurun_out/bench_s2c1.err; tail -c 3000 gpurun_out/bench_s2c1.json; tail -5 gpurun_out/bench_s2c1.err
echo "== cfg5 sweep"; timeout 600 python tools/sweep.py cfg5 > gpurun_out/sweep_cfg5_s2c1.json 2> gpurun_out/sweep_cfg5_s2c1.err; tail -5 gpurun_out/sweep_cfg5_s2c1.err
