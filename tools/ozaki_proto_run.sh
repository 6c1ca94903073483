mkdir -p gpurun_out
P=tools/ozaki_glm_proto
{
timeout 60 $P check 8192 64 3 5 || echo "{\"rc\": $?}"
timeout 60 $P time 250000 4096 37 2 || echo "{\"rc\": $?}"
timeout 60 $P time 250000 4096 37 5 || echo "{\"rc\": $?}"
timeout 120 $P time 1000000 4096 37 5 || echo "{\"rc\": $?}"
} > gpurun_out/ozaki_proto6.jsonl 2>&1
cat gpurun_out/ozaki_proto6.jsonl
