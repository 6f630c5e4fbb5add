#!/bin/bash
# round-2 call 10 (8 GPUs): headline config with (a) the NVLS gather + measured fused/library selector, (b) selector only
mkdir -p gpurun_out/r2c10
O=gpurun_out/r2c10
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MLB200_AG_NVLS=1 timeout 200 $TR --master-port 29590 bench.py --gpus 8 --steps 4 --warmup 3 --no_e2e > $O/llama7b_tp8_nvls_ours.json 2> $O/llama7b_tp8_nvls_ours.err; echo "nvls rc=$? $(grep -o '"value": [0-9.]*' $O/llama7b_tp8_nvls_ours.json | head -1) $(grep -o '"exposed_tp_collective_ms_per_step": {[^}]*}' $O/llama7b_tp8_nvls_ours.json)"
timeout 200 $TR --master-port 29591 bench.py --gpus 8 --steps 4 --warmup 3 --no_e2e > $O/llama7b_tp8_selector_ours.json 2> $O/llama7b_tp8_selector_ours.err; echo "selector rc=$? $(grep -o '"value": [0-9.]*' $O/llama7b_tp8_selector_ours.json | head -1) $(grep -o '"exposed_tp_collective_ms_per_step": {[^}]*}' $O/llama7b_tp8_selector_ours.json)"
MLB200_AG_NVLS=1 MLB200_FUSED_TP_FORCE=1 timeout 200 $TR --master-port 29592 bench.py --gpus 8 --steps 4 --warmup 3 --no_e2e > $O/llama7b_tp8_nvls_force_ours.json 2> $O/llama7b_tp8_nvls_force_ours.err; echo "nvls+force rc=$? $(grep -o '"value": [0-9.]*' $O/llama7b_tp8_nvls_force_ours.json | head -1)"
