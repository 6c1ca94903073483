#!/bin/bash
mkdir -p gpurun_out
for cx in 32 16 8; do for as in 0 1; do
  echo -n "sleepstudy 16384 x 150 CX=$cx async=$as: "; SMC_NUTS_POST_CX=$cx SMC_NUTS_ASYNC=$as timeout 600 python tools/nuts_baseline_run.py sleepstudy 16384 150 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g' % d['leapfrogs_per_s'], d['rounds'], round(d['active_lane_fraction'],3))"
done; done
for cx in 32 8; do
  echo -n "hierarchical 16384 x 150 CX=$cx async=1: "; SMC_NUTS_POST_CX=$cx SMC_NUTS_ASYNC=1 timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 150 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g' % d['leapfrogs_per_s'], d['rounds'], round(d['active_lane_fraction'],3))"
done
echo "== sampler tests with CX=8"; SMC_NUTS_POST_CX=8 timeout 900 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -3
