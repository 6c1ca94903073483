#!/bin/bash
mkdir -p gpurun_out
export SMC_GLM_I8_CHECK=1 SMC_NO_GRAPHS=1
for f in 64 3; do
  echo "== flush $f"; SMC_GLM_I8_FLUSH=$f timeout 200 python tools/i8_stress.py 400000 1024 6 2>&1 | tail -2
done
