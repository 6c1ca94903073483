#!/bin/bash
# session 2, call 43: launch lists of a simpleNUTS round with the final code (sleepstudy in lockstep, hierarchical in lockstep;
# the same commands as s2_call31.sh, for a before / after of the evaluation's launch list)
mkdir -p gpurun_out
for model in cfg4 cfg4h; do
  SMC_NUTS_ASYNC=0 NUTS_STEPS=24 SMC_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 1500 --csv --log-file gpurun_out/launches_${model}_16k_v4.csv python tools/prof_cfg.py $model 16384 > gpurun_out/ncu_${model}_16k_v4.log 2>&1; tail -1 gpurun_out/ncu_${model}_16k_v4.log
done
