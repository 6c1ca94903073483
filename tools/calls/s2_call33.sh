#!/bin/bash
mkdir -p gpurun_out
echo "== gpu suite with SMC_NO_GRAPHS=1"; ( time SMC_NO_GRAPHS=1 timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1
echo "== gpu suite with SMC_NO_POOL=1"; ( time SMC_NO_POOL=1 timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1
