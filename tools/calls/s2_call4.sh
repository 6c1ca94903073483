#!/bin/bash
# session 2, call 4: int8 kernel with optimistic slicing (no CTA barrier between epilogue 1 and the slicing): parity, time, profile;
# launch list of cfg3 with 2 chains (where does a leapfrog's time go); DRAM traffic of one full-size launch
mkdir -p gpurun_out
echo "== i8 quick"; timeout 120 python tools/i8_quick.py; SMC_GLM_I8_BETA8=1 timeout 120 python tools/i8_quick.py
echo "== i8 parity tests"; timeout 900 python -m pytest tests/test_zz_gpu_i8.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -5
echo "== i8 at configs[1]"; timeout 300 python tools/i8_bench_extra.py 10 3
echo "== i8 profile (mid size)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm_i8 -s 2 -c 1 -o gpurun_out/prof_i8_v13 python tools/i8_prof.py > gpurun_out/ncu_i8_v13.log 2>&1; tail -2 gpurun_out/ncu_i8_v13.log
echo "== i8 traffic (full size)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_glm_i8 -s 2 -c 1 --csv --log-file gpurun_out/i8_traffic_full.csv python tools/i8_prof.py 1000000 4096 > gpurun_out/ncu_i8_traffic.log 2>&1; tail -2 gpurun_out/ncu_i8_traffic.log; tail -5 gpurun_out/i8_traffic_full.csv
echo "== cfg3 2 chains launch list"
SMC_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3_c2.csv python tools/prof_cfg.py cfg3 2 > gpurun_out/ncu_cfg3_c2.log 2>&1; tail -2 gpurun_out/ncu_cfg3_c2.log
