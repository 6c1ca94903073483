mkdir -p gpurun_out
{
timeout 60 tools/ozaki_glm_proto_m check 8192 64 3 7 || echo "{\"rc\": $?}"
timeout 60 tools/ozaki_glm_proto_m time 250000 4096 37 7 || echo "{\"rc\": $?}"
timeout 60 tools/ozaki_glm_proto time 250000 4096 37 7 || echo "{\"rc\": $?}"
} > gpurun_out/ozaki_proto9.jsonl 2>&1
cat gpurun_out/ozaki_proto9.jsonl
