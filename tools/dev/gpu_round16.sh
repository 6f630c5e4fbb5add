#!/bin/bash
mkdir -p gpurun_out
MLB200_GEMM2_DEBUG=1 MLB200_GEMM_2CTA=1 timeout 200 python tools/dev/gemm2_debug.py > gpurun_out/gemm2_debug2.txt 2>&1; tail -10 gpurun_out/gemm2_debug2.txt
out=gpurun_out/gemm_checks_2cta.jsonl; : > $out
for c in "nt 4096 4096 4096 t" "nt 4096 12288 4096 t" "nt 4096 22016 4096 t" "nn 4096 4096 11008 t" "nn 4096 4096 4096 t" "tn_acc 22016 4096 4096 t" "tn_acc 4096 4096 4096 t" "nt 520 264 200" "nt 4096 32000 4096 t"; do
  MLB200_GEMM_2CTA=1 timeout 120 python tools/dev/gpu_check_gemm.py $c >> $out 2>/dev/null || echo "{\"case\": \"$c\", \"failed\": $?}" >> $out
done
python - <<'PY'
import json
for l in open("gpurun_out/gemm_checks_2cta.jsonl"):
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ("case","M","N","K","ok","rel_err","tflops","cublas_tflops","failed")})
PY
timeout 600 python bench.py --steps 3 --warmup 3 --no_e2e --graph 1 > gpurun_out/bench_7b_r16_graph.json 2> gpurun_out/bench_7b_r16_graph.err; tail -1 gpurun_out/bench_7b_r16_graph.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')})" || tail -5 gpurun_out/bench_7b_r16_graph.err
