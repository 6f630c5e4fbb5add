#!/bin/bash
# ncu --set full capture of every hot kernel, one launch each (one GPU: ncu replays a kernel ~40 times).
#   tools/profiling/ncu_kernels.sh [outdir]      -> <outdir>/ncu_<kernel>.ncu-rep (+ .log)
# then, anywhere:  python tools/profiling/ncu_report.py <outdir>/*.ncu-rep > profiles/ncu_summary.md
# Numbers printed by a command running under ncu are never benchmark values.
out=${1:-gpurun_out}
mkdir -p "$out"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 300 $NCU -k "regex:$regex" -s "$skip" -c 1 -o "$out/ncu_$name" "$@" > "$out/ncu_$name.log" 2>&1
  echo "$name: exit $? $(ls -la "$out/ncu_$name.ncu-rep" 2>/dev/null | awk '{print $5}') bytes"
}
G="python tools/profiling/gemm_check.py"
A="python tools/profiling/attn_check.py"
MLB200_GEMM_2CTA=0 cap gemm_1cta_nt      gemm_bf16_kernel       2 $G nt 4096 22016 4096 t
cap gemm_2cta_nt      gemm_bf16_2cta_kernel  2 $G nt 4096 22016 4096 t
cap gemm_2cta_nn      gemm_bf16_2cta_kernel  2 $G nn 4096 4096 11008 t
cap gemm_2cta_wgrad   gemm_bf16_2cta_kernel  2 $G tn_acc 22016 4096 4096 t
cap attn_fwd2         attn_fwd2_kernel       1 $A 1 4096 32 32 none t
cap attn_bwd_dkdv     attn_bwd_dkdv_kernel   1 $A 1 4096 32 32 none t
cap attn_bwd_dq       attn_bwd_dq_kernel     1 $A 1 4096 32 32 none t
E="python tools/profiling/ew_drive.py"
cap rmsnorm_fwd       norm_fwd_kernel        2 $E
cap rmsnorm_bwd       norm_bwd_kernel        1 $E
cap swiglu_fwd        glu_fwd_kernel         2 $E
cap swiglu_bwd        glu_bwd_kernel         1 $E
cap rope_qkv          rope_qkv_kernel        1 $E
cap ce_stats          ce_stats_kernel        1 $E
cap adamw_flat        adamw_flat_kernel      1 $E
cap embedding_bwd     embedding_bwd_kernel   1 $E
cap bias_dropout_add  bias_dropout_add_kernel 1 $E
# the fused GEMM+collective kernels with ONE virtual rank (ncu serialises kernels, so peers cannot answer): instruction
# mix / tensor pipe of the same code that moves tiles through peer pointers
L="python tools/profiling/fused_loopback_drive.py"
cap fused_ag_gemm     gemm_bf16_2cta_kernel  1 $L ag
cap fused_gemm_rs     gemm_bf16_2cta_kernel  1 $L rs
