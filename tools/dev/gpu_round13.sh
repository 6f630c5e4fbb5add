#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/attn_checks6.jsonl; : > $out
for c in "1 128 1 1 none c" "1 1024 2 2 256 c" "2 1024 4 1 none c" "1 2048 8 2 512 c" "1 4096 32 32 none t" "1 4096 32 8 none t" "4 2048 32 32 none t" "1 8192 32 32 4096 t"; do
  timeout 120 python tools/dev/attn_check.py $c >> $out 2> gpurun_out/attn6_err_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
for c in "2 512 2 1 none" "1 1024 1 2 256"; do
  timeout 120 python tools/dev/attn_packed_check.py $c >> $out 2> gpurun_out/attn6_err_packed_$(echo $c | tr ' ' '_').txt || echo "{\"case\": \"packed $c\", \"failed\": $?}" >> $out
done
cat $out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_bf16_kernel -s 2 -c 1 -f -o gpurun_out/ncu_gemm_1cta python tools/dev/gpu_check_gemm.py nt 4096 22016 4096 t > gpurun_out/ncu_gemm_1cta.log 2>&1
MLB200_GEMM_2CTA=1 timeout 300 $NCU -k regex:gemm_bf16_2cta -s 2 -c 1 -f -o gpurun_out/ncu_gemm_2cta python tools/dev/gpu_check_gemm.py nt 4096 22016 4096 t > gpurun_out/ncu_gemm_2cta.log 2>&1
timeout 300 $NCU -k regex:attn_fwd2_kernel -s 1 -c 1 -f -o gpurun_out/ncu_attn_fwd2_v4 python tools/dev/attn_check.py 1 4096 32 32 none t > gpurun_out/ncu_attn_fwd2_v4.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_gemm_1cta.log gpurun_out/ncu_gemm_2cta.log
timeout 600 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_r13.json 2> gpurun_out/bench_7b_r13.err; tail -1 gpurun_out/bench_7b_r13.json | cut -c1-400
