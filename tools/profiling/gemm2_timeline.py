"""Timeline counters of the 2-CTA GEMM (MLB200_GEMM2_DEBUG=1 MLB200_GEMM_2CTA=1)."""
import sys, torch
sys.path.insert(0, ".")
from megatron_llm_b200 import ops
from megatron_llm_b200.ops import _ext
M, N, K = 4096, 22016, 4096
a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16); b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
out = ops.gemm_nt(a, b)
for _ in range(3): ops.gemm_nt(a, b, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); ops.gemm_nt(a, b, out=out); e.record(); torch.cuda.synchronize()
print("ms", s.elapsed_time(e), "TFLOPS", 2 * M * N * K / s.elapsed_time(e) / 1e9)
d = _ext.load().gemm2_debug()[:148].double()
names = ["mma_wait_full", "mma_wait_tmem_empty", "mma_total", "prod_wait_empty", "prod_total", "epi_wait_tmem_full", "epi_total", "tiles"]
lead, peer = d[0::2], d[1::2]
for i, n in enumerate(names):
    print(f"{n:22s} leader mean {lead[:, i].mean():12.0f} max {lead[:, i].max():12.0f} | peer mean {peer[:, i].mean():12.0f} max {peer[:, i].max():12.0f}")
