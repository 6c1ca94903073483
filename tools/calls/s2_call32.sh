#!/bin/bash
mkdir -p gpurun_out
export SMC_NUTS_ASYNC=0 NUTS_STEPS=24
timeout 900 ncu --set full --clock-control none -k regex:k_nuts_post -s 1500 -c 1 -o /tmp/nuts_post_final python tools/prof_cfg.py cfg4 16384 > gpurun_out/ncu_nuts_post_final.log 2>&1; tail -1 gpurun_out/ncu_nuts_post_final.log
ncu -i /tmp/nuts_post_final.ncu-rep --page raw --csv > gpurun_out/nuts_post_final_raw.csv 2>/dev/null
