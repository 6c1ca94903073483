#!/bin/bash
mkdir -p gpurun_out
echo "== new posterior test"; timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
echo "== sampler + parity tests with SMC_NUTS_ASYNC=1"; SMC_NUTS_ASYNC=1 timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "== sampler tests with SMC_NUTS_DEVICE_LOOP=1"; SMC_NUTS_DEVICE_LOOP=1 timeout 900 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -3
