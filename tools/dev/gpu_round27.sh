#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 300 python -m pytest tests/test_tp_model_gpu.py -m gpu -x -q --basetemp=gpurun_out/r27_tmp > gpurun_out/r27_tp_model.log 2>&1
echo "exit $?"; tail -25 gpurun_out/r27_tp_model.log | cut -c1-1500
cat gpurun_out/r27_tmp/*/fast.json; echo; cat gpurun_out/r27_tmp/*/plain.json
timeout 120 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "cuda_graph" > gpurun_out/r27_graph_test.log 2>&1
echo "graph test exit $?"; tail -3 gpurun_out/r27_graph_test.log
