#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -x --timeout 600 > gpurun_out/pytest_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_model.log
tail -30 gpurun_out/pytest_model.log
timeout 900 python -X faulthandler bench.py --steps 3 --warmup 3 > gpurun_out/bench_7b.json 2> gpurun_out/bench_7b.err; echo "7b rc=$?"
tail -3 gpurun_out/bench_7b.json; grep -v "^$" gpurun_out/bench_7b.err | tail -30
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/ref_7b.json 2> gpurun_out/ref_7b.err; echo "ref 7b rc=$?"
tail -2 gpurun_out/ref_7b.json; grep -v "Warning\|warn\|^  " gpurun_out/ref_7b.err | tail -15
