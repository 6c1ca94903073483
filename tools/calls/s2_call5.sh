#!/bin/bash
# session 2, call 5: int8 kernel v13 (optimistic slicing, per-stage barB) parity / time / profile / traffic; cfg1 k_glm with the tail
# columns of product 2 on DFMAs (A/B against SMC_GLM_NO_TAIL=1) + parity; cfg3 C=2 launch list
mkdir -p gpurun_out
echo "== i8 quick"; timeout 120 python tools/i8_quick.py; SMC_GLM_I8_BETA8=1 timeout 120 python tools/i8_quick.py
echo "== i8 parity tests"; timeout 900 python -m pytest tests/test_zz_gpu_i8.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -5
echo "== i8 at configs[1]"; timeout 300 python tools/i8_bench_extra.py 10 3
echo "== cfg1 sweep: tail on / off"; timeout 300 python tools/sweep.py cfg1 > gpurun_out/sweep_cfg1_tail.json 2>gpurun_out/sweep_cfg1_tail.err; grep -E "chains|leapfrog_chain|glm_tflops" gpurun_out/sweep_cfg1_tail.json | paste - - - | tail -4
SMC_GLM_NO_TAIL=1 timeout 300 python tools/sweep.py cfg1 > gpurun_out/sweep_cfg1_notail.json 2>gpurun_out/sweep_cfg1_notail.err; grep -E "chains|leapfrog_chain|glm_tflops" gpurun_out/sweep_cfg1_notail.json | paste - - - | tail -4
echo "== parity of the narrow GLM shapes"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_cfg3.py -m gpu -x -q 2>&1 | tail -4
echo "== i8 profile (mid size)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm_i8 -s 2 -c 1 -o gpurun_out/prof_i8_v13 python tools/i8_prof.py > gpurun_out/ncu_i8_v13.log 2>&1; tail -2 gpurun_out/ncu_i8_v13.log
echo "== i8 traffic (full size)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_glm_i8 -s 2 -c 1 --csv --log-file gpurun_out/i8_traffic_full.csv python tools/i8_prof.py 1000000 4096 > gpurun_out/ncu_i8_traffic.log 2>&1; tail -2 gpurun_out/ncu_i8_traffic.log; tail -5 gpurun_out/i8_traffic_full.csv
echo "== cfg1 profile"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm -s 4 -c 1 -o gpurun_out/prof_cfg1_tail python tools/prof_cfg.py cfg1 > gpurun_out/ncu_cfg1_tail.log 2>&1; tail -2 gpurun_out/ncu_cfg1_tail.log
