#!/bin/bash
# round-2 call 3 (1 GPU): stream debug for the graph test, then the whole GPU suite (no -x)
mkdir -p gpurun_out/r2c3
O=gpurun_out/r2c3
export MASTER_ADDR=127.0.0.1
echo "== stream debug (cpu init)"; timeout 300 python tools/profiling/debug_graph_stream.py > $O/dbg_cpuinit.log 2>&1; grep STREAM $O/dbg_cpuinit.log
echo "== stream debug (gpu init)"; timeout 300 python tools/profiling/debug_graph_stream.py gpuinit > $O/dbg_gpuinit.log 2>&1; grep STREAM $O/dbg_gpuinit.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 $O/pytest_gpu.log
