#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/attn_checks.jsonl; : > $out
for c in "1 128 1 1 none c" "1 256 2 2 none c" "2 512 4 2 none c" "1 1024 2 2 256 c" "1 4096 32 32 none t" "1 4096 32 8 none t"; do
  timeout 180 python tools/dev/attn_check.py $c >> $out 2> gpurun_out/attn_err_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
cat $out
tail -5 gpurun_out/attn_err_1_128_1_1_none_c.txt
