#!/bin/bash
mkdir -p gpurun_out
export SMC_NUTS_ASYNC=1 NUTS_STEPS=24
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nuts_post -s 1500 -c 1 -o /tmp/nuts_post_async python tools/prof_cfg.py cfg4 16384 > gpurun_out/ncu_nuts_post.log 2>&1; tail -1 gpurun_out/ncu_nuts_post.log
ncu -i /tmp/nuts_post_async.ncu-rep --page raw --csv > gpurun_out/nuts_post_async_raw.csv 2>/dev/null
ncu -i /tmp/nuts_post_async.ncu-rep --page source --csv --print-source sass > gpurun_out/nuts_post_async_src.csv 2>/dev/null
du -sh gpurun_out
