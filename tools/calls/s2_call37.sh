#!/bin/bash
# session 2, call 37: int8 kernel v16 (optimistic slicing of v13 + the phase split of v14)
mkdir -p gpurun_out
echo "== stress"
for f in 64 3 1; do
  SMC_GLM_I8_CHECK=1 SMC_NO_GRAPHS=1 SMC_GLM_I8_FLUSH=$f timeout 200 python tools/i8_stress.py 400000 1024 6 2>&1 | tail -1
done
echo "== i8 quick"; timeout 120 python tools/i8_quick.py
echo "== i8 at configs[1]"; timeout 300 python tools/i8_bench_extra.py 10 3
echo "== i8 parity tests"; timeout 900 python -m pytest tests/test_zz_gpu_i8.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -5
