#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dev/debug_nan.py 1024 > gpurun_out/debug_nan_1024.log 2>&1; echo "rc=$?" >> gpurun_out/debug_nan_1024.log
grep -v "^  \|\.\.\.\.\." gpurun_out/debug_nan_1024.log | tail -20
MLB200_DISABLE_KERNELS=1 timeout 300 python tools/dev/debug_nan.py 1024 > gpurun_out/debug_nan_1024_nok.log 2>&1
grep -v "^  \|\.\.\.\.\." gpurun_out/debug_nan_1024_nok.log | tail -8
timeout 600 python -X faulthandler bench.py --steps 2 --warmup 1 --no_e2e > gpurun_out/bench_7b.json 2> gpurun_out/bench_7b.err; echo "7b rc=$?"
tail -3 gpurun_out/bench_7b.json; grep -v "^$" gpurun_out/bench_7b.err | tail -40
timeout 600 python bench.py --impl reference --model llama2-tiny --steps 3 --warmup 2 > gpurun_out/ref_tiny.json 2> gpurun_out/ref_tiny.err; echo "ref tiny rc=$?"
tail -2 gpurun_out/ref_tiny.json; tail -25 gpurun_out/ref_tiny.err
