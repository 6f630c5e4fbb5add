#!/bin/bash
# round-2 call 14 (4 GPUs): exposed DP-reduction / PP-p2p accounting on the small TP x DP and TP x PP layouts
mkdir -p gpurun_out/r2c14
O=gpurun_out/r2c14
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
MLB200_FUSED_TP=1 timeout 60 $TR --master-port 29620 bench.py --gpus 4 --steps 3 --warmup 3 --no_e2e --model mistral-tiny --tp 2 --dist_opt > $O/mistral_tiny.json 2> $O/mistral_tiny.err; echo "dp rc=$? $(grep -o '"value": [0-9.]*' $O/mistral_tiny.json | head -1) $(grep -o '"exposed_dp_reduce_ms_per_step": [0-9.a-z]*' $O/mistral_tiny.json)"
timeout 60 $TR --master-port 29621 bench.py --gpus 4 --steps 3 --warmup 3 --no_e2e --model falcon-tiny --tp 2 --pp 2 --global_batch 16 > $O/falcon_tiny.json 2> $O/falcon_tiny.err; echo "pp rc=$? $(grep -o '"value": [0-9.]*' $O/falcon_tiny.json | head -1) $(grep -o '"exposed_pp_p2p_ms_per_step": [0-9.a-z]*' $O/falcon_tiny.json)"
tail -3 $O/*.err | cut -c1-200
