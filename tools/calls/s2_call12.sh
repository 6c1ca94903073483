#!/bin/bash
# session 2, call 12: int8 kernel v14 (tensor bursts aligned with the integer-only phases of the epilogues)
mkdir -p gpurun_out
echo "== stress"
for f in 64 3 1; do
  SMC_GLM_I8_CHECK=1 SMC_NO_GRAPHS=1 SMC_GLM_I8_FLUSH=$f timeout 200 python tools/i8_stress.py 400000 1024 6 2>&1 | tail -1
done
echo "== i8 quick"; timeout 120 python tools/i8_quick.py; SMC_GLM_I8_BETA8=1 timeout 120 python tools/i8_quick.py
echo "== i8 parity tests"; timeout 900 python -m pytest tests/test_zz_gpu_i8.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -5
echo "== i8 at configs[1]"; timeout 300 python tools/i8_bench_extra.py 10 3
echo "== i8 profile (mid size)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glm_i8 -s 2 -c 1 -o /tmp/i8_v14 python tools/i8_prof.py > gpurun_out/ncu_i8_v14.log 2>&1; tail -1 gpurun_out/ncu_i8_v14.log
ncu -i /tmp/i8_v14.ncu-rep --page raw --csv > gpurun_out/i8_v14_raw.csv 2>/dev/null
ncu -i /tmp/i8_v14.ncu-rep --page source --csv --print-source sass > gpurun_out/i8_v14_src.csv 2>/dev/null
du -sh gpurun_out
