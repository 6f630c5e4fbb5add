#!/bin/bash
# each case isolated in its own process + timeout so a hang cannot take the box down
mkdir -p gpurun_out
out=gpurun_out/gemm_checks.jsonl; : > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
for c in "nt 256 256 128" "nt 128 256 64" "nt 4096 4096 4096 t" "nt 4096 1536 4096 t" "nt 4096 22016 4096 t" "nt 520 264 200" \
         "nn 256 256 128" "nn 4096 4096 4096 t" "nn 4096 4096 11008 t" \
         "tn 256 256 128" "tn 4096 4096 4096 t" "tn_acc 256 512 384" "tn_acc 22016 4096 4096 t" "nt 8192 8192 8192 t"; do
  timeout 120 python tools/profiling/gemm_check.py $c >> $out 2>gpurun_out/gemm_err_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
cat $out
