#!/bin/bash
mkdir -p gpurun_out
for mode in 1 0; do
  export SMC_NUTS_ASYNC=$mode
  NUTS_STEPS=24 SMC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 1500 --csv --log-file gpurun_out/launches_cfg4_16k_async$mode.csv python tools/prof_cfg.py cfg4 16384 > gpurun_out/ncu_cfg4_16k_$mode.log 2>&1; tail -1 gpurun_out/ncu_cfg4_16k_$mode.log
done
