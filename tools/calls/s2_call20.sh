#!/bin/bash
# suspend-time hint of the int8 kernel's mbarrier waits: 2000 (library), 500, 100 ns
mkdir -p gpurun_out
cd simplemcmc.jl_b200; cp libsimplemcmc_cuda.so /tmp/lib_orig.so; cd ..
for v in orig w500 w100; do
  if [ $v = orig ]; then cp /tmp/lib_orig.so simplemcmc.jl_b200/libsimplemcmc_cuda.so; else cp simplemcmc.jl_b200/libsimplemcmc_cuda_$v.so simplemcmc.jl_b200/libsimplemcmc_cuda.so; fi
  echo "== $v"; timeout 300 python tools/i8_bench_extra.py 10 3 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['glm_ms_per_launch'], d['value'], d['relerr_grad_vs_long_double'])"
done
cp /tmp/lib_orig.so simplemcmc.jl_b200/libsimplemcmc_cuda.so
