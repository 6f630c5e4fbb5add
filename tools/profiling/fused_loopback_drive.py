"""One virtual rank (world = 1 loopback) through the fused AG->GEMM / GEMM->RS kernels at a Llama-2-7B layer shape, for
ncu (which serialises kernels: real peers could not answer).  python tools/profiling/fused_loopback_drive.py ag|rs"""
import sys
import torch
sys.path.insert(0, ".")
from megatron_llm_b200.parallel.symm import TPCommunicator
kind = sys.argv[1] if len(sys.argv) > 1 else "ag"
comm = TPCommunicator.loopback_group(1, max_rows_per_rank=4096, max_k=11008, max_n=11008, num_comm_ctas=4, sms=148)[0]
bf = torch.bfloat16
if kind == "ag":
    x = torch.randn(4096, 4096, device="cuda", dtype=bf)
    w = torch.randn(11008, 4096, device="cuda", dtype=bf) * 0.02
    for _ in range(3):
        comm.ag_gemm(x, w, False)
else:
    a = torch.randn(4096, 11008, device="cuda", dtype=bf)
    w = torch.randn(4096, 11008, device="cuda", dtype=bf) * 0.02
    for _ in range(3):
        comm.gemm_rs(a, w, False)
torch.cuda.synchronize()
assert comm.error_flag() == 0
