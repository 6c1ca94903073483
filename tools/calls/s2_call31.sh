#!/bin/bash
mkdir -p gpurun_out
for model in cfg4 cfg4h; do
  SMC_NUTS_ASYNC=0 NUTS_STEPS=24 SMC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 2000 --csv --log-file gpurun_out/launches_${model}_16k_r02.csv python tools/prof_cfg.py $model 16384 > gpurun_out/ncu_${model}_16k.log 2>&1; tail -1 gpurun_out/ncu_${model}_16k.log
done
