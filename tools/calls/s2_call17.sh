#!/bin/bash
mkdir -p gpurun_out
for mode in async lockstep; do
  if [ $mode = lockstep ]; then export SMC_NUTS_LOCKSTEP=1; else unset SMC_NUTS_LOCKSTEP; fi
  echo "== $mode: plain"; NUTS_STEPS=30 timeout 300 python tools/prof_cfg.py cfg4 1024
  echo "== $mode: launch list (kernels of rounds in the middle of the run)"
  NUTS_STEPS=30 SMC_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 1500 --csv --log-file gpurun_out/launches_cfg4_${mode}.csv python tools/prof_cfg.py cfg4 1024 > gpurun_out/ncu_cfg4_${mode}.log 2>&1; tail -1 gpurun_out/ncu_cfg4_${mode}.log
done
