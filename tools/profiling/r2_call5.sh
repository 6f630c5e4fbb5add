#!/bin/bash
# round-2 call 5 (4 GPUs): the three non-headline BASELINE layouts on SMALL models, both arms -- shakes out TP x DP (one
# fused TP communicator per group + peer-memory DP reduction + fused ZeRO gather), TP x PP (1F1B) and TP + recompute on
# hardware before the 8-GPU runs
mkdir -p gpurun_out/r2c5
O=gpurun_out/r2c5
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() {  # name, extra bench args...
  local name=$1; shift
  for impl in ours reference; do
    timeout 400 $TR --master-port 29560 bench.py --impl $impl --gpus 4 --steps 3 --warmup 3 "$@" > $O/${name}_$impl.json 2> $O/${name}_$impl.err
    echo "$name $impl rc=$? $(tail -c 600 $O/${name}_$impl.json | head -c 600)"
  done
}
run mistral_tiny_tp2dp2_zero1 --model mistral-tiny --tp 2 --dist_opt
MLB200_FUSED_TP=1 run mistral_tiny_tp2dp2_zero1_fusedtp --model mistral-tiny --tp 2 --dist_opt --no_e2e
run falcon_tiny_tp2pp2 --model falcon-tiny --tp 2 --pp 2 --global_batch 16
run llama_tiny_tp4_recompute --model llama2-tiny --recompute
grep -l "Traceback" $O/*.err | head
echo "== multi-GPU tests (fused TP kernels incl. all-reduce, DP kernels peer / NVLS, TP invariance 1/2/4)"
MLB200_TEST_RECORD=$O/tp_invariance.json timeout 900 python -m pytest tests/test_fused_comm_gpu.py tests/test_tp_model_gpu.py -m gpu -q --timeout 600 > $O/pytest_multigpu.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_multigpu.log; cat $O/tp_invariance.json 2>/dev/null; echo
echo "== DP reduction kernels vs NCCL (1 GB fp32 bucket, 4 GPUs)"
timeout 300 $TR --master-port 29561 tools/profiling/dp_bench.py 1024 > $O/dp_bench_n4.jsonl 2> $O/dp_bench.err; cat $O/dp_bench_n4.jsonl; tail -3 $O/dp_bench.err
