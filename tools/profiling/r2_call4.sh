#!/bin/bash
# round-2 call 4 (2 GPUs): whole GPU suite incl. multi-GPU tests + TP invariance, N=2 bench (fused / nccl / reference),
# fused per-pair times from graph replays, multicast probe
mkdir -p gpurun_out/r2c4
O=gpurun_out/r2c4
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== multicast probe"; timeout 120 $TR --master-port 29550 tools/profiling/mc_probe.py > $O/mc_probe.json 2> $O/mc_probe.err; cat $O/mc_probe.json
echo "== pytest -m gpu (all, 2 GPUs)"
MLB200_TEST_RECORD=$O/tp_invariance.json timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu.log; cat $O/tp_invariance.json 2>/dev/null | head -c 1500; echo
echo "== bench N=2 ours (fused + graph, auto micro-batch)"
timeout 600 $TR --master-port 29551 bench.py --gpus 2 --steps 4 --warmup 3 > $O/bench_n2_ours.json 2> $O/bench_n2_ours.err; echo "rc=$?"; tail -c 1800 $O/bench_n2_ours.json
echo "== bench N=2 ours (NCCL collectives + graph)"
MLB200_FUSED_TP=0 timeout 600 $TR --master-port 29552 bench.py --gpus 2 --steps 4 --warmup 3 --no_e2e > $O/bench_n2_ours_nccl.json 2> $O/bench_n2_ours_nccl.err; echo "rc=$?"; tail -c 900 $O/bench_n2_ours_nccl.json
echo "== bench N=2 ours micro-batch 1 (fused + graph)"
timeout 600 $TR --master-port 29553 bench.py --gpus 2 --steps 4 --warmup 3 --micro_batch 1 --no_e2e > $O/bench_n2_ours_mb1.json 2> $O/bench_n2_ours_mb1.err; echo "rc=$?"; tail -c 900 $O/bench_n2_ours_mb1.json
echo "== bench N=2 reference"
timeout 900 $TR --master-port 29554 bench.py --impl reference --gpus 2 --steps 4 --warmup 3 > $O/bench_n2_ref.json 2> $O/bench_n2_ref.err; echo "rc=$?"; tail -c 1200 $O/bench_n2_ref.json
echo "== fused per-pair (graph replay timing): seq 4096 and 8192 (micro-batch 2)"
for seq in 4096 8192; do
  timeout 200 $TR --master-port 29555 tools/profiling/fused_bench.py 4096 11008 $seq > $O/fused_bench_n2_s$seq.jsonl 2> $O/fused_bench_s$seq.err; echo "seq $seq rc=$? ($(grep -c '^{' $O/fused_bench_n2_s$seq.jsonl) shapes)"
done
