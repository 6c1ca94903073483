#!/bin/bash
# session 2, call 9: whole gpu suite; cfg1 with head + tail columns on DFMAs; cfg5 sweep (no-gradient tape on the compiled shape); bench
mkdir -p gpurun_out
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1
echo "== cfg1 sweep"; timeout 300 python tools/sweep.py cfg1 > gpurun_out/sweep_cfg1_r02.json 2>gpurun_out/sweep_cfg1_r02.err; grep -E "chains|leapfrog_chain|glm_tflops" gpurun_out/sweep_cfg1_r02.json | paste - - - | tail -4
echo "== cfg5 sweep"; timeout 600 python tools/sweep.py cfg5 > gpurun_out/sweep_cfg5_r02.json 2> gpurun_out/sweep_cfg5_r02.err; tail -3 gpurun_out/sweep_cfg5_r02.err
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err; tail -c 1500 gpurun_out/bench_1gpu_r02.json; tail -3 gpurun_out/bench_1gpu_r02.err
echo "== cfg1 profile"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm -s 4 -c 1 -o gpurun_out/prof_cfg1_tail python tools/prof_cfg.py cfg1 > gpurun_out/ncu_cfg1_tail.log 2>&1; tail -2 gpurun_out/ncu_cfg1_tail.log
echo "== cfg3 C=64 profile (Bernoulli family on the DMMA kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_glm -s 4 -c 1 -o gpurun_out/prof_cfg3_c64 python tools/prof_cfg.py cfg3 64 > gpurun_out/ncu_cfg3_c64.log 2>&1; tail -2 gpurun_out/ncu_cfg3_c64.log
echo "== i8 profile (v12 + fma)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm_i8 -s 2 -c 1 -o gpurun_out/prof_i8_v12b python tools/i8_prof.py > gpurun_out/ncu_i8_v12b.log 2>&1; tail -2 gpurun_out/ncu_i8_v12b.log
