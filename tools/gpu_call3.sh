#!/bin/bash
mkdir -p gpurun_out
{
  echo "== i8 quick (N=3000, M=50, 70 chains)"
  SMC_GLM_I8=1 timeout 120 python tools/i8_quick.py
  SMC_GLM_I8=1 SMC_GLM_I8_BETA8=1 timeout 120 python tools/i8_quick.py
  echo "== i8 opt-in parity"
  SMC_TEST_I8=1 timeout 900 python -m pytest tests/test_zz_gpu_i8_optin.py -m gpu -x -q 2>&1 | tail -5
  echo "== i8 at configs[1]"
  SMC_GLM_I8=1 timeout 300 python tools/i8_bench_extra.py 10 3
  SMC_GLM_I8=1 SMC_GLM_I8_BETA8=1 timeout 300 python tools/i8_bench_extra.py 10 3
  echo "== GLM-related suites with the switch set"
  SMC_GLM_I8=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_samplers.py -m gpu -q 2>&1 | tail -8
} > gpurun_out/call3.log 2>&1
tail -30 gpurun_out/call3.log
