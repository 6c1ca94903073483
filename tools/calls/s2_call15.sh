#!/bin/bash
# session 2, call 15: where a NUTS round's time goes at 16 384 chains (launch list), packing below 4096 chains
mkdir -p gpurun_out
echo "== cfg4 sweep with packing forced at 1024 chains"
SMC_COMPACT=1 timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_compact.json 2>/dev/null; grep -E "chains|leapfrogs_per_s|active_lane" gpurun_out/sweep_cfg4_compact.json | paste - - - | tail -4
echo "== launch lists"
SMC_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file gpurun_out/launches_cfg4_sleep.csv python tools/prof_cfg.py cfg4 > gpurun_out/ncu_cfg4.log 2>&1; tail -1 gpurun_out/ncu_cfg4.log
SMC_NO_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file gpurun_out/launches_cfg4_hier.csv python tools/prof_cfg.py cfg4h > gpurun_out/ncu_cfg4h.log 2>&1; tail -1 gpurun_out/ncu_cfg4h.log
du -sh gpurun_out
