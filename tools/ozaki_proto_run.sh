#!/bin/bash
# The measured variants of tools/ozaki_glm_proto.cu (profiles/ozaki_glm_proto_v*_r01.jsonl): parity against the host long-double
# reference, then time at 250k rows and at the full configs[1] size.   gpurun --timeout 300 -- 'bash tools/ozaki_proto_run.sh'
mkdir -p gpurun_out
P=tools/ozaki_glm_proto
{
for v in 0 1 2 5 7; do timeout 60 $P check 8192 64 3 $v || echo "{\"rc\": $?, \"variant\": $v}"; done
for v in 0 1 2 5 7; do timeout 60 $P time 250000 4096 37 $v || echo "{\"rc\": $?, \"variant\": $v}"; done
timeout 120 $P time 1000000 4096 37 7 || echo "{\"rc\": $?}"
} > gpurun_out/ozaki_proto.jsonl 2>&1
cat gpurun_out/ozaki_proto.jsonl
