mkdir -p gpurun_out
P=tools/ozaki_glm_proto
{
timeout 60 $P check 8192 64 3 || echo "{\"rc\": $?}"
timeout 90 $P time 250000 4096 37 || echo "{\"rc\": $?}"
} > gpurun_out/ozaki_proto5.jsonl 2>&1
cat gpurun_out/ozaki_proto5.jsonl


