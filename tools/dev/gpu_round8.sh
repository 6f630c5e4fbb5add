#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/attn_checks2.jsonl; : > $out
( time python -c "import torch; torch.zeros(1).cuda(); import megatron_llm_b200.ops._ext as e; e.load(); print('warm')" ) > gpurun_out/warm.txt 2>&1
for c in "1 128 1 1 none c" "1 256 2 2 none c" "1 1024 2 2 256 c" "2 1024 4 1 none c"; do
  CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/dev/attn_check.py $c >> $out 2> gpurun_out/attn2_err_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
cat $out
for f in gpurun_out/attn2_err_*; do echo "== $f"; tail -25 $f | cut -c1-300; done
cat gpurun_out/warm.txt | tail -5
