#!/bin/bash
mkdir -p gpurun_out
for ly in 8 1; do for as in 0 1; do
  echo "== sleepstudy 16384 x 200, SMC_NUTS_LY=$ly SMC_NUTS_ASYNC=$as"; SMC_NUTS_LY=$ly SMC_NUTS_ASYNC=$as timeout 600 python tools/nuts_baseline_run.py sleepstudy 16384 200 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['leapfrogs_per_s'], d['rounds'], d['active_lane_fraction'], d['device_s'])"
done; done
echo "== hierarchical 16384 x 200 LY=1 async"; SMC_NUTS_LY=1 SMC_NUTS_ASYNC=1 timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 200 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['leapfrogs_per_s'], d['rounds'], d['active_lane_fraction'], d['device_s'])"
echo "== hierarchical 16384 x 200 LY=8 async"; SMC_NUTS_LY=8 SMC_NUTS_ASYNC=1 timeout 600 python tools/nuts_baseline_run.py hierarchical 16384 200 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['leapfrogs_per_s'], d['rounds'], d['active_lane_fraction'], d['device_s'])"
echo "== sampler tests with LY=1"; SMC_NUTS_LY=1 timeout 900 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -3
