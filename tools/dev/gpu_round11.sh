#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/attn_checks5.jsonl; : > $out
for c in "1 128 1 1 none c" "1 1024 2 2 256 c" "2 1024 4 1 none c" "1 2048 8 2 512 c" "1 4096 32 32 none t" "1 4096 32 8 none t"; do
  timeout 120 python tools/dev/attn_check.py $c >> $out 2> gpurun_out/attn5_err_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
for c in "2 512 2 1 none" "1 1024 2 4 none" "1 1024 1 2 256"; do
  timeout 120 python tools/dev/attn_packed_check.py $c >> $out 2> gpurun_out/attn5_err_packed_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"packed $c\", \"failed\": $?}" >> $out
done
cat $out
for f in gpurun_out/attn5_err_*; do if grep -q "Error\|error\|Fatal" $f; then echo "== $f"; grep -v "^  File" $f | tail -8 | cut -c1-300; fi; done
timeout 600 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_r11.json 2> gpurun_out/bench_7b_r11.err; tail -1 gpurun_out/bench_7b_r11.json | cut -c1-400
