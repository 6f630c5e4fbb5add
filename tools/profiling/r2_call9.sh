#!/bin/bash
# round-2 call 9 (2 GPUs): NVLS all-gather transport -- numerics (eager + graph replay), per-pair times, short bench
mkdir -p gpurun_out/r2c9
O=gpurun_out/r2c9
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_fused_comm_gpu.py -m gpu -q --timeout 300 -k "nvls or match_nccl" > $O/pytest_nvls.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_nvls.log | cut -c1-300
MLB200_AG_NVLS=1 timeout 200 $TR --master-port 29580 tools/profiling/fused_bench.py 4096 11008 8192 > $O/fused_bench_n2_nvls_s8192.jsonl 2> $O/fused_bench_nvls.err; echo "fused_bench rc=$?"; grep '^{' $O/fused_bench_n2_nvls_s8192.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['kind'], r['what'], 'gemm', round(r['gemm_ms']*1e3), 'nccl+gemm', round(r['nccl_plus_gemm_ms']*1e3), 'fused', round(r['fused_ms']*1e3))"
tail -3 $O/fused_bench_nvls.err | cut -c1-300
MLB200_AG_NVLS=1 timeout 300 $TR --master-port 29581 bench.py --gpus 2 --steps 3 --warmup 3 --no_e2e > $O/bench_n2_nvls.json 2> $O/bench_n2_nvls.err; echo "bench rc=$?"; tail -c 700 $O/bench_n2_nvls.json
