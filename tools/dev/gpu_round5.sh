#!/bin/bash
# 2-GPU round: fused-kernel correctness, then TP=2 bench with the unfused NCCL path and with the fused kernels
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_fused_comm_gpu.py -q -x --timeout 300 > gpurun_out/pytest_fused.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fused.log
grep -v "^  File\|^    " gpurun_out/pytest_fused.log | tail -40
for mode in 0 1; do
  MLB200_FUSED_TP=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 \
      bench.py --gpus 2 --steps 3 --warmup 3 --no_e2e > gpurun_out/bench_7b_tp2_fused$mode.json 2> gpurun_out/bench_7b_tp2_fused$mode.err; echo "tp2 fused=$mode rc=$?"
  tail -2 gpurun_out/bench_7b_tp2_fused$mode.json | cut -c1-600; grep -v "Warning\|warn\|^  \|^$" gpurun_out/bench_7b_tp2_fused$mode.err | tail -12
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 \
      bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/ref_7b_tp2.json 2> gpurun_out/ref_7b_tp2.err; echo "ref tp2 rc=$?"
tail -2 gpurun_out/ref_7b_tp2.json | cut -c1-600; grep -v "Warning\|warn\|^  \|^$" gpurun_out/ref_7b_tp2.err | tail -8
