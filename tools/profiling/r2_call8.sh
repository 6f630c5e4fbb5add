#!/bin/bash
# round-2 call 8 (8 GPUs): the headline config and BASELINE configs 3-5, ours first (all four), then the reference arm,
# then the NCCL-collective variant of ours.  Every run writes its own JSON; a run that dies leaves its .err.
mkdir -p gpurun_out/r2c8
O=gpurun_out/r2c8
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
one() {  # name, timeout, impl, args...
  local name=$1 to=$2 impl=$3; shift 3
  local t0=$(date +%s)
  timeout $to $TR --master-port 29570 bench.py --impl $impl --gpus 8 "$@" > $O/${name}_$impl.json 2> $O/${name}_$impl.err
  echo "$name $impl rc=$? $(( $(date +%s) - t0 ))s $(grep -o '"value": [0-9.]*' $O/${name}_$impl.json | head -1) $(grep -o '"peak_mem_gb": [0-9.]*' $O/${name}_$impl.json | head -1) $(grep -o '"unavailable.*' $O/${name}_$impl.json | head -c 200)"
}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_before.txt 2>&1
one llama7b_tp8          240 ours      --steps 4 --warmup 3
MLB200_FUSED_TP=1 one mistral7b_tp2dp4     240 ours      --model mistral-7b --tp 2 --dist_opt --steps 3 --warmup 3
MLB200_FUSED_TP=1 one falcon40b_tp4pp2     300 ours      --model falcon-40b --tp 4 --pp 2 --global_batch 16 --steps 2 --warmup 3 --no_e2e
if ! grep -q '"value"' $O/falcon40b_tp4pp2_ours.json; then
  mv $O/falcon40b_tp4pp2_ours.err $O/falcon40b_tp4pp2_ours_fused_failed.err
  MLB200_FUSED_TP=0 one falcon40b_tp4pp2   300 ours      --model falcon-40b --tp 4 --pp 2 --global_batch 16 --steps 2 --warmup 3 --no_e2e
fi
one llama70b_tp8_recomp  300 ours      --model llama2-70b --recompute --micro_batch 4 --steps 2 --warmup 3 --no_e2e
one llama7b_tp8          300 reference --steps 4 --warmup 3
one mistral7b_tp2dp4     300 reference --model mistral-7b --tp 2 --dist_opt --steps 3 --warmup 3
one falcon40b_tp4pp2     360 reference --model falcon-40b --tp 4 --pp 2 --global_batch 16 --steps 2 --warmup 3
one llama70b_tp8_recomp  360 reference --model llama2-70b --recompute --micro_batch 4 --steps 2 --warmup 3
MLB200_FUSED_TP=0 one llama7b_tp8_nccl 240 ours --steps 4 --warmup 3 --no_e2e
timeout 150 $TR --master-port 29571 tools/profiling/fused_bench.py 4096 11008 32768 > $O/fused_bench_n8_mb8.jsonl 2> $O/fused_bench_n8.err; echo "fused_bench rc=$? $(grep -c '^{' $O/fused_bench_n8_mb8.jsonl) shapes"
