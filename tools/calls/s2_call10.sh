#!/bin/bash
# session 2, call 10: conditional-graph probe; NUTS device loop tests; whole gpu suite; cfg1 / cfg4 / cfg5 sweeps; bench; profiles
# (ncu reports stay in /tmp on the box: only their CSV pages come back, gpurun_out is limited to 64 MiB)
mkdir -p gpurun_out
echo "== conditional graph probe"; timeout 60 ./tools/cond_graph_probe; timeout 60 ./tools/cond_graph_probe child
echo "== NUTS tests"; timeout 900 python -m pytest tests/test_gpu_samplers.py -m gpu -x -q 2>&1 | tail -6
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1
echo "== cfg1 sweep"; timeout 300 python tools/sweep.py cfg1 > gpurun_out/sweep_cfg1_r02.json 2>gpurun_out/sweep_cfg1_r02.err; grep -E "chains|leapfrog_chain|glm_tflops" gpurun_out/sweep_cfg1_r02.json | paste - - - | tail -4
echo "== cfg4 sweep"; timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_r02.json 2>gpurun_out/sweep_cfg4_r02.err; grep -E "chains|leapfrogs_per_s|launches" gpurun_out/sweep_cfg4_r02.json | paste - - - | tail -4
SMC_NUTS_HOST_LOOP=1 timeout 600 python tools/sweep.py cfg4 > gpurun_out/sweep_cfg4_hostloop_r02.json 2>gpurun_out/sweep_cfg4_hostloop_r02.err; grep -E "chains|leapfrogs_per_s|launches" gpurun_out/sweep_cfg4_hostloop_r02.json | paste - - - | tail -4
echo "== cfg5 sweep"; timeout 600 python tools/sweep.py cfg5 > gpurun_out/sweep_cfg5_r02.json 2> gpurun_out/sweep_cfg5_r02.err; tail -3 gpurun_out/sweep_cfg5_r02.err
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_1gpu_r02.json 2> gpurun_out/bench_1gpu_r02.err; tail -c 1500 gpurun_out/bench_1gpu_r02.json; tail -3 gpurun_out/bench_1gpu_r02.err
prof() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/${name}_src.csv 2>/dev/null
}
echo "== profiles"
prof cfg1_tail_r02 k_glm 4 python tools/prof_cfg.py cfg1
prof cfg3_c64_r02 k_glm 4 python tools/prof_cfg.py cfg3 64
prof i8_v12b_r02 k_glm_i8 2 python tools/i8_prof.py
prof cfg5_c64_r02 k_affnorm 4 python tools/prof_cfg.py cfg5 64 hmc
du -sh gpurun_out
