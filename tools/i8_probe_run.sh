mkdir -p gpurun_out
P=tools/i8_umma_probe
{
for mode in kk mk km; do for sw in 0 1; do timeout 20 $P check $mode 64 $sw || echo "{\"rc\": $?, \"mode\": \"$mode\", \"swap\": $sw}"; done; done
timeout 20 $P check kk 256 0; timeout 20 $P check kk 128 0; timeout 20 $P check km 128 0; timeout 20 $P check mk 256 0
for n in 64 128 256; do timeout 30 $P rate $n 20000 || echo "{\"rc\": $?}"; done
} > gpurun_out/i8_probe1.jsonl 2>&1
cat gpurun_out/i8_probe1.jsonl
