#!/bin/bash
mkdir -p gpurun_out
MLB200_GEMM2_DEBUG=1 MLB200_GEMM_2CTA=1 timeout 200 python tools/dev/gemm2_debug.py > gpurun_out/gemm2_debug.txt 2>&1; cat gpurun_out/gemm2_debug.txt | tail -12
timeout 600 python bench.py --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_r15.json 2> gpurun_out/bench_7b_r15.err; tail -1 gpurun_out/bench_7b_r15.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','host_enqueue_ms_per_step')})"
