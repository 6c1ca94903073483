#!/bin/bash
# session 2, call 13: the committed tree -- whole gpu suite, bench line, int8 kernel (v14) profile + DRAM traffic of a full-size launch
mkdir -p gpurun_out
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err; tail -c 600 gpurun_out/bench_1gpu_r02.json; tail -3 gpurun_out/bench_1gpu_r02.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference_arm_r02.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_arm_r02.json
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_r02.log 2>&1; tail -1 gpurun_out/ncu_launches_r02.log | cut -c1-200
echo "== i8 profile (v14, mid size) + traffic (full size)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm_i8 -s 2 -c 1 -o /tmp/i8_v14 python tools/i8_prof.py > gpurun_out/ncu_i8_v14.log 2>&1; tail -1 gpurun_out/ncu_i8_v14.log
ncu -i /tmp/i8_v14.ncu-rep --page raw --csv > gpurun_out/i8_v14_raw.csv 2>/dev/null
ncu -i /tmp/i8_v14.ncu-rep --page source --csv --print-source sass > gpurun_out/i8_v14_src.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_glm_i8 -s 2 -c 1 --csv --log-file gpurun_out/i8_traffic_full.csv python tools/i8_prof.py 1000000 4096 > gpurun_out/ncu_i8_traffic.log 2>&1; tail -1 gpurun_out/ncu_i8_traffic.log
du -sh gpurun_out
