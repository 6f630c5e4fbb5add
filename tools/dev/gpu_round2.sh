#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -8 gpurun_out/smoke.log
timeout 300 python bench.py --model llama2-tiny --steps 3 --warmup 2 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "tiny rc=$?"
tail -3 gpurun_out/bench_tiny.json; tail -5 gpurun_out/bench_tiny.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_7b.json 2> gpurun_out/bench_7b.err; echo "7b rc=$?"
tail -3 gpurun_out/bench_7b.json; tail -15 gpurun_out/bench_7b.err
