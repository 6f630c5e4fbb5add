#!/bin/bash
# Validation of the head_dim-64 attention instantiations (passed on a B200 in round 2; the kernels are on by default):
# numerics of fwd / dQ / dK / dV vs the fp32 reference for MHA, GQA, MQA, sliding window, separate and packed QKV, and
# timing vs the FA-2 library on the Falcon-7B / GPT shapes.  Each case in its own process under a timeout.
mkdir -p gpurun_out; out=gpurun_out/attn_hd64_checks.jsonl; : > $out
for c in "1 128 1 1 none c 64" "1 256 2 2 none c 64" "2 512 4 2 none c 64" "2 512 8 1 none c 64" "1 1024 2 2 256 c 64" \
         "1 2048 71 1 none t 64" "1 4096 16 16 none t 64"; do
  timeout 180 python tools/profiling/attn_check.py $c >> $out 2> gpurun_out/attn_hd64_err_$(echo $c | tr ' ' '_').txt \
    || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
for c in "1 256 2 1 none 64" "2 512 2 4 none 64" "1 1024 1 8 256 64"; do
  timeout 180 python tools/profiling/attn_packed_check.py $c >> $out 2>> gpurun_out/attn_hd64_packed_err.txt \
    || echo "{\"packed_case\": \"$c\", \"failed\": $?}" >> $out
done
cat $out
