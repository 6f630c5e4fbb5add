#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/gemm_checks_2cta_v2.jsonl; : > $out
for c in "nt 256 256 128" "nt 520 264 200" "nt 4096 22016 4096 t" "nn 4096 4096 11008 t" "nn 264 520 136" "tn 328 264 520" "tn_acc 256 512 384" "tn_acc 520 264 200" "tn_acc 22016 4096 4096 t" "tn_acc 4096 4096 4096 t" "tn_acc 4096 11008 4096 t"; do
  MLB200_GEMM_2CTA=1 timeout 120 python tools/dev/gpu_check_gemm.py $c >> $out 2>gpurun_out/gemm2_err.txt || { echo "{\"case\": \"$c\", \"failed\": $?}" >> $out; tail -3 gpurun_out/gemm2_err.txt; }
done
python - <<'PY'
import json
for l in open("gpurun_out/gemm_checks_2cta_v2.jsonl"):
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k in ("case","M","N","K","ok","rel_err","tflops","cublas_tflops","failed")})
PY
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_7b_r19.json 2> gpurun_out/bench_7b_r19.err; tail -1 gpurun_out/bench_7b_r19.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','e2e')})" || tail -5 gpurun_out/bench_7b_r19.err
