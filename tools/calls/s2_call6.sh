#!/bin/bash
# session 2, call 6: stress of the int8 kernel's pipeline protocol with wait-site codes
mkdir -p gpurun_out
export SMC_GLM_I8_CHECK=1 SMC_NO_GRAPHS=1
for f in 64 1 2 3 7; do
  echo "== flush $f"; SMC_GLM_I8_FLUSH=$f timeout 200 python tools/i8_stress.py 400000 1024 6 2>&1 | tail -2
done
echo "== flush 64, full size tile count"; timeout 300 python tools/i8_stress.py 1000000 4096 4 2>&1 | tail -2
