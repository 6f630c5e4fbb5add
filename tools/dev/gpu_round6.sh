#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dev/gemm_shapes_bench.py > gpurun_out/gemm_shapes.jsonl 2> gpurun_out/gemm_shapes.err; echo "gemm rc=$?"
cat gpurun_out/gemm_shapes.jsonl; tail -5 gpurun_out/gemm_shapes.err
timeout 600 python tools/dev/profile_step.py 8 gpurun_out/profile_step_8l.txt > gpurun_out/profile.log 2>&1; echo "profile rc=$?"
cat gpurun_out/profile_step_8l.txt | head -50
