#!/bin/bash
mkdir -p gpurun_out
SMC_NUTS_ASYNC=1 NUTS_STEPS=24 SMC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 2000 --csv --log-file gpurun_out/launches_cfg4_16k_async_final.csv python tools/prof_cfg.py cfg4 16384 > gpurun_out/ncu_cfg4_16k_async_final.log 2>&1; tail -1 gpurun_out/ncu_cfg4_16k_async_final.log
