#!/bin/bash
mkdir -p gpurun_out
echo "== cfg1 sweep: 128-chain tile at three CTAs per SM"; SMC_GLM_TC128_3=1 timeout 300 python tools/sweep.py cfg1 > gpurun_out/sweep_cfg1_tc128.json 2>/dev/null; grep -E "chains|leapfrog_chain|glm_tflops" gpurun_out/sweep_cfg1_tc128.json | paste - - - | tail -3
echo "== cfg1 sweep: default (256-chain tile, two CTAs per SM)"; timeout 300 python tools/sweep.py cfg1 > gpurun_out/sweep_cfg1_default.json 2>/dev/null; grep -E "chains|leapfrog_chain|glm_tflops" gpurun_out/sweep_cfg1_default.json | paste - - - | tail -3
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_cfg3.py -m gpu -x -q 2>&1 | tail -3
