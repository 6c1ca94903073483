#!/bin/bash
mkdir -p gpurun_out
echo "== sampler + parity + fuzz tests with SMC_NUTS_ASYNC=1 (mode-comparison tests excluded)"; SMC_NUTS_ASYNC=1 timeout 900 python -m pytest tests/test_gpu_samplers.py tests/test_gpu_parity.py -m gpu -q -k "not invisible" 2>&1 | tail -4
echo "== sampler tests with SMC_NUTS_DEVICE_LOOP=1 (mode-comparison tests excluded)"; SMC_NUTS_DEVICE_LOOP=1 timeout 900 python -m pytest tests/test_gpu_samplers.py -m gpu -q -k "not invisible" 2>&1 | tail -4
