#!/bin/bash
mkdir -p gpurun_out
{
  echo "== i8 quick"
  SMC_GLM_I8=1 timeout 120 python tools/i8_quick.py
  SMC_GLM_I8=1 SMC_GLM_I8_BETA8=1 timeout 120 python tools/i8_quick.py
  echo "== i8 opt-in parity"
  SMC_TEST_I8=1 timeout 900 python -m pytest tests/test_zz_gpu_i8_optin.py -m gpu -x -q 2>&1 | tail -5
  echo "== i8 at configs[1]"
  SMC_GLM_I8=1 timeout 300 python tools/i8_bench_extra.py 10 3
  echo "== AFFNORM + solvers + cfg3 full size + samplers"
  timeout 1200 python -m pytest tests/test_gpu_affnorm.py tests/test_gpu_scale_cfg3.py tests/test_gpu_samplers.py tests/test_gpu_solvers.py -m gpu -q 2>&1 | tail -15
  echo "== GLM-related suites with the switch set"
  SMC_GLM_I8=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q 2>&1 | tail -8
  echo "== cfg5 sweep"
  timeout 900 python tools/sweep.py cfg5 > gpurun_out/sweep_cfg5_r02.json 2> gpurun_out/sweep_cfg5_r02.err; tail -3 gpurun_out/sweep_cfg5_r02.err
  echo "== cfg5 launch list + i8 profile"
  SMC_NO_GRAPHS=1 timeout 300 python tools/cfg5_prof.py && \
  SMC_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_affnorm --csv --log-file gpurun_out/launches_cfg5_r02.csv python tools/cfg5_prof.py > gpurun_out/ncu_cfg5.log 2>&1; tail -2 gpurun_out/ncu_cfg5.log
  SMC_GLM_I8=1 timeout 120 python tools/i8_prof.py && \
  SMC_GLM_I8=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm_i8 -s 2 -c 1 -o gpurun_out/prof_i8_v12 python tools/i8_prof.py > gpurun_out/ncu_i8_v12.log 2>&1; tail -3 gpurun_out/ncu_i8_v12.log
} > gpurun_out/call6.log 2>&1
tail -60 gpurun_out/call6.log
