#!/bin/bash
# k_nuts_post at 233 registers runs one block per SM: minimum blocks per SM 1 (library), 2, 3, 4
mkdir -p gpurun_out
cp simplemcmc.jl_b200/libsimplemcmc_cuda.so /tmp/lib_orig.so
for v in orig b2 b3 b4; do
  if [ $v = orig ]; then cp /tmp/lib_orig.so simplemcmc.jl_b200/libsimplemcmc_cuda.so; else cp simplemcmc.jl_b200/libsimplemcmc_cuda_$v.so simplemcmc.jl_b200/libsimplemcmc_cuda.so; fi
  for as in 0 1; do
    echo -n "$v sleepstudy 16384 x 150 async=$as: "; SMC_NUTS_ASYNC=$as timeout 600 python tools/nuts_baseline_run.py sleepstudy 16384 150 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g' % d['leapfrogs_per_s'], d['rounds'], round(d['active_lane_fraction'],3))"
  done
  echo -n "$v hierarchical 16384 x 150 async=1: "; SMC_NUTS_ASYNC=1 timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 150 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g' % d['leapfrogs_per_s'], d['rounds'], round(d['active_lane_fraction'],3))"
done
cp /tmp/lib_orig.so simplemcmc.jl_b200/libsimplemcmc_cuda.so
