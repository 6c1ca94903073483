#!/bin/bash
mkdir -p gpurun_out
{
  echo "== AFFNORM parity"
  timeout 900 python -m pytest tests/test_gpu_affnorm.py -m gpu -q 2>&1 | tail -25
  echo "== cfg5 sweep lines"
  timeout 600 python tools/sweep.py cfg5 2>&1 | tail -12
  echo "== i8 profile"
  SMC_GLM_I8=1 timeout 120 python tools/i8_prof.py && \
  SMC_GLM_I8=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm_i8 -s 2 -c 1 -o gpurun_out/prof_i8_v10 python tools/i8_prof.py > gpurun_out/ncu_i8_v10.log 2>&1; tail -3 gpurun_out/ncu_i8_v10.log
} > gpurun_out/call4.log 2>&1
tail -50 gpurun_out/call4.log
