#!/bin/bash
# round-2 call 2 (1 GPU): whole GPU suite (incl. loopback fused kernels, fp32-oracle parity, attention grads),
# head_dim-64 attention validation, short N=1 bench
mkdir -p gpurun_out/r2c2
O=gpurun_out/r2c2
export MASTER_ADDR=127.0.0.1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu.log
echo "== experimental (hd64 attention, fp16)"
MLB200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -k "head_dim_64 or fp16" > $O/pytest_experimental.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_experimental.log
echo "== bench N=1"
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"; tail -c 1500 $O/bench_n1.json
