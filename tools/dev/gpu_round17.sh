#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/dev/gemm_shapes_bench.py > gpurun_out/gemm_shapes_v2.jsonl 2> gpurun_out/gemm_shapes_v2.err; python - <<'PY'
import json
for l in open("gpurun_out/gemm_shapes_v2.jsonl"):
    d=json.loads(l); print(d['kind'], d['M'], d['N'], d['K'], 'cta1', d.get('cta1_tflops'), 'cta2', d.get('cta2_tflops'), 'cublas', d['cublas_tflops'], d.get('cta2_relerr'))
PY
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_r17.json 2> gpurun_out/bench_7b_r17.err; tail -1 gpurun_out/bench_7b_r17.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')})" || tail -5 gpurun_out/bench_7b_r17.err
timeout 600 python bench.py --steps 3 --warmup 3 --no_e2e --graph 1 > gpurun_out/bench_7b_r17_graph.json 2> gpurun_out/bench_7b_r17_graph.err; tail -1 gpurun_out/bench_7b_r17_graph.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')})" || (grep -v "^\s*$" gpurun_out/bench_7b_r17_graph.err | grep -B12 "Error\|error" | grep -v OMP | tail -30 | cut -c1-250)
