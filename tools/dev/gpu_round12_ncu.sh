#!/bin/bash
# ncu --set full captures of the hot kernels (one GPU; every kernel replayed ~40x, so few launches each)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_bf16_kernel -s 3 -c 1 -f -o gpurun_out/ncu_gemm_1cta python tools/dev/gpu_check_gemm.py nt 4096 22016 4096 > gpurun_out/ncu_gemm_1cta.log 2>&1
MLB200_GEMM_2CTA=1 timeout 300 $NCU -k regex:gemm_bf16_2cta -s 3 -c 1 -f -o gpurun_out/ncu_gemm_2cta python tools/dev/gpu_check_gemm.py nt 4096 22016 4096 > gpurun_out/ncu_gemm_2cta.log 2>&1
timeout 300 $NCU -k regex:attn_fwd2_kernel -s 1 -c 1 -f -o gpurun_out/ncu_attn_fwd2 python tools/dev/attn_check.py 1 4096 32 32 none t > gpurun_out/ncu_attn_fwd2.log 2>&1
timeout 300 $NCU -k regex:attn_bwd_dkdv -s 1 -c 1 -f -o gpurun_out/ncu_attn_dkdv python tools/dev/attn_check.py 1 4096 32 32 none t > gpurun_out/ncu_attn_dkdv.log 2>&1
timeout 300 $NCU -k regex:attn_bwd_dq -s 1 -c 1 -f -o gpurun_out/ncu_attn_dq python tools/dev/attn_check.py 1 4096 32 32 none t > gpurun_out/ncu_attn_dq.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_gemm_2cta.log
