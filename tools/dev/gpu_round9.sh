#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/attn_checks3.jsonl; : > $out
export MLB200_ATTN_DEBUG=1
for c in "1 128 1 1 none c" "2 512 4 2 none c"; do
  timeout 100 python tools/dev/attn_check.py $c >> $out 2> gpurun_out/attn3_err_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
CUDA_LAUNCH_BLOCKING=1 timeout 100 python tools/dev/attn_check.py 2 512 4 2 none c >> $out 2> gpurun_out/attn3_err_blocking.txt || echo "{\"case\": \"blocking\", \"failed\": $?}" >> $out
timeout 200 cuda-gdb -batch -ex run -ex bt -ex "info threads" --args python tools/dev/attn_check.py 1 128 1 1 none c > gpurun_out/attn3_gdb.txt 2>&1
cat $out
for f in gpurun_out/attn3_err_*; do echo "== $f"; grep -v "^  File\|^$\|Extension modules" $f | tail -12 | cut -c1-200; done
echo "== gdb"; tail -40 gpurun_out/attn3_gdb.txt | cut -c1-250
