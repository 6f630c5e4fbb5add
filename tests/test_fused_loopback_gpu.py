"""The fused GEMM+collective kernels and the peer-memory DP reduction on ONE GPU: ``world`` virtual ranks share the
device (parallel/symm.py::LoopbackWorld), every rank's kernel runs on its own stream with 1/world of the SMs, and the
ranks talk to each other through the same flag / epoch / receive-slot protocol they use over NVLink -- the "peer"
pointers are simply other allocations of this process.  Oracle: fp32 PyTorch matmul + explicit gather / reduce."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_ranks(world, fn):
    """fn(rank) on ``world`` streams (all launches enqueued before anything is awaited); returns the per-rank results."""
    streams = [torch.cuda.Stream() for _ in range(world)]
    cur = torch.cuda.current_stream()
    out = []
    for r in range(world):
        streams[r].wait_stream(cur)
        with torch.cuda.stream(streams[r]):
            out.append(fn(r))
    for st in streams:
        cur.wait_stream(st)
    return out


@pytest.mark.parametrize("world,m", [(2, 256), (2, 128), (4, 256), (8, 256)])
def test_loopback_ag_gemm_and_gemm_rs(world, m):
    """m = 256 takes the 2-CTA (cta_group::2) fused kernels, m = 128 the 1-CTA ones; several calls in a row so the
    epochs, the receive-slot parities and the puller acks are exercised."""
    from megatron_llm_b200.parallel.symm import TPCommunicator
    K, N = 512, 768
    comms = TPCommunicator.loopback_group(world, max_rows_per_rank=512, max_k=2048, max_n=2048, num_comm_ctas=4)
    torch.manual_seed(1)
    for it in range(3):
        xs = [torch.randn(m, K, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
        ws = [torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.05 for _ in range(world)]
        wts = [torch.randn(K, N, device=DEV, dtype=torch.bfloat16) * 0.05 for _ in range(world)]
        full = torch.cat(xs, 0)
        res = _run_ranks(world, lambda r: comms[r].ag_gemm(xs[r], ws[r], False))
        res_t = _run_ranks(world, lambda r: comms[r].ag_gemm(xs[r], wts[r], True))
        torch.cuda.synchronize()
        for r in range(world):
            out, gathered = res[r]
            assert torch.equal(gathered, full), f"it {it} rank {r}: gathered mismatch"
            ref = full.float() @ ws[r].float().t()
            assert (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max(), f"it {it} rank {r}: ag_gemm"
            ref_t = full.float() @ wts[r].float()
            assert (res_t[r][0].float() - ref_t).abs().max() <= 2e-2 * ref_t.abs().max(), f"it {it} rank {r}: ag_gemm(T)"
        # GEMM -> reduce-scatter: rank r contributes a_r @ w_r^T, and receives rows [r*m, (r+1)*m) of the sum
        a = [torch.randn(world * m, K, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
        got = _run_ranks(world, lambda r: comms[r].gemm_rs(a[r], ws[r], False))
        got_t = _run_ranks(world, lambda r: comms[r].gemm_rs(a[r], wts[r], True))
        torch.cuda.synchronize()
        total = sum(a[r].float() @ ws[r].float().t() for r in range(world))
        total_t = sum(a[r].float() @ wts[r].float() for r in range(world))
        for r in range(world):
            ref = total[r * m:(r + 1) * m]
            assert (got[r].float() - ref).abs().max() <= 3e-2 * ref.abs().max(), f"it {it} rank {r}: gemm_rs"
            ref = total_t[r * m:(r + 1) * m]
            assert (got_t[r].float() - ref).abs().max() <= 3e-2 * ref.abs().max(), f"it {it} rank {r}: gemm_rs(T)"
    for c in comms:
        assert c.error_flag() == 0, "a spin-wait timed out"


def test_loopback_tolerates_rank_skew():
    """Ranks that are early, late or a whole call behind (device-side sleeps of different lengths before each launch):
    READY / ACK / ARRIVED / FREE handshakes must still produce the right answer without a timeout."""
    from megatron_llm_b200.parallel.symm import TPCommunicator
    world, m, K, N = 2, 256, 512, 768
    comms = TPCommunicator.loopback_group(world, max_rows_per_rank=512, max_k=2048, max_n=2048, num_comm_ctas=4)
    rnd = random.Random(7)
    torch.manual_seed(2)
    w = [torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.05 for _ in range(world)]
    results = []
    for it in range(8):
        xs = [torch.randn(m, K, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
        a = [torch.randn(world * m, K, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
        delays = [rnd.randrange(0, 2_000_000) for _ in range(2 * world)]

        def rank_step(r):
            torch.cuda._sleep(delays[2 * r])
            o = comms[r].ag_gemm(xs[r], w[r], False)
            torch.cuda._sleep(delays[2 * r + 1])
            return o, comms[r].gemm_rs(a[r], w[r], False)
        results.append((xs, a, _run_ranks(world, rank_step)))       # no host sync between iterations
    torch.cuda.synchronize()
    for it, (xs, a, res) in enumerate(results):
        full = torch.cat(xs, 0)
        total = sum(a[r].float() @ w[r].float().t() for r in range(world))
        for r in range(world):
            (out, gathered), got = res[r]
            assert torch.equal(gathered, full), f"it {it} rank {r}: gathered mismatch"
            ref = full.float() @ w[r].float().t()
            assert (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max(), f"it {it} rank {r}: ag_gemm"
            ref = total[r * m:(r + 1) * m]
            assert (got.float() - ref).abs().max() <= 3e-2 * ref.abs().max(), f"it {it} rank {r}: gemm_rs"
    for c in comms:
        assert c.error_flag() == 0, "a spin-wait timed out"


@pytest.mark.parametrize("world,reduce_scatter", [(2, False), (4, False), (4, True)])
def test_loopback_dp_reduce(world, reduce_scatter):
    from megatron_llm_b200.parallel.symm import DPCommunicator
    n = 1 << 18
    comms = DPCommunicator.loopback_group(world, n)
    for it in range(3):
        torch.manual_seed(10 + it)
        g = [torch.randn(n, device=DEV) for _ in range(world)]
        ref = sum(g) / world
        for r in range(world):
            comms[r].buffer.copy_(g[r])
        torch.cuda.synchronize()
        half = n // 2
        hs = _run_ranks(world, lambda r: (comms[r].reduce_bucket(comms[r].buffer[:half], 0, n, reduce_scatter),
                                          comms[r].reduce_bucket(comms[r].buffer[half:], half, n, reduce_scatter)))
        for h1, h2 in hs:
            h1.wait(); h2.wait()
        torch.cuda.synchronize()
        for r in range(world):
            if reduce_scatter:      # rank r owns slice r of each bucket
                sl = half // world
                for base in (0, half):
                    assert torch.allclose(comms[r].buffer[base + r * sl: base + (r + 1) * sl],
                                          ref[base + r * sl: base + (r + 1) * sl], atol=1e-5), f"it {it} rank {r}"
            else:
                assert torch.allclose(comms[r].buffer, ref, atol=1e-5), f"it {it} rank {r}"


@pytest.mark.parametrize("world,m", [(2, 256), (4, 256), (2, 128)])
def test_loopback_gemm_all_reduce(world, m):
    """GEMM -> all-reduce (non-sequence-parallel Row forward / Column dgrad): every rank ends with the full [M, N] sum.
    Interleaved with plain reduce-scatter calls: both share the receive slots, parities and epochs."""
    from megatron_llm_b200.parallel.symm import TPCommunicator
    K, N = 512, 768
    M = world * m
    comms = TPCommunicator.loopback_group(world, max_rows_per_rank=512, max_k=2048, max_n=2048, num_comm_ctas=4,
                                          all_reduce_n=1024)
    torch.manual_seed(3)
    for it in range(4):
        a = [torch.randn(M, K, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
        ws = [torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.05 for _ in range(world)]
        wts = [torch.randn(K, N, device=DEV, dtype=torch.bfloat16) * 0.05 for _ in range(world)]
        got = _run_ranks(world, lambda r: comms[r].gemm_ar(a[r], ws[r], False))
        if it % 2:
            _run_ranks(world, lambda r: comms[r].gemm_rs(a[r], ws[r], False))       # shifts the slot parity
        got_t = _run_ranks(world, lambda r: comms[r].gemm_ar(a[r], wts[r], True, keep=False))
        torch.cuda.synchronize()
        total = sum(a[r].float() @ ws[r].float().t() for r in range(world))
        total_t = sum(a[r].float() @ wts[r].float() for r in range(world))
        for r in range(world):
            assert got[r].shape == (M, N)
            assert (got[r].float() - total).abs().max() <= 3e-2 * total.abs().max(), f"it {it} rank {r}: gemm_ar"
            assert (got_t[r].float() - total_t).abs().max() <= 3e-2 * total_t.abs().max(), f"it {it} rank {r}: gemm_ar(T)"
        for r in range(1, world):                    # every rank holds bit-identical results
            assert torch.equal(got[r], got[0])
    for c in comms:
        assert c.error_flag() == 0, "a spin-wait timed out"


def test_loopback_zero1_param_gather_barrier():
    """ZeRO-1 fused cast + parameter all-gather: every virtual rank's AdamW kernel stores its shard of the updated
    16-bit weights into all ranks' parameter buffers, then the ranks meet in the peer barrier."""
    from megatron_llm_b200.ops import _ext
    from megatron_llm_b200.parallel.symm import DPCommunicator
    world, n = 4, 1 << 16
    comms = DPCommunicator.loopback_group(world, n)
    for c in comms:
        c.attach_param_buffer(n, torch.bfloat16)
    mod = _ext.load()
    shard = n // world
    torch.manual_seed(21)
    master = torch.randn(n, device=DEV)
    grads = torch.randn(n, device=DEV)
    seg_start = torch.tensor([0, n], device=DEV, dtype=torch.int64)
    seg_wd = torch.zeros(1, device=DEV)
    coef = torch.ones(1, device=DEV)
    inf = torch.zeros(1, device=DEV, dtype=torch.int32)
    state = [(master[r * shard:(r + 1) * shard].clone(), torch.zeros(shard, device=DEV), torch.zeros(shard, device=DEV))
             for r in range(world)]

    def step(r):
        p, m, v = state[r]
        mod.adamw_flat(p, grads[r * shard:(r + 1) * shard].contiguous(), m, v, comms[r].pbuf[r * shard:(r + 1) * shard],
                       r * shard, seg_start, seg_wd, None, 1e-2, 0.9, 0.95, 1e-8, 0.1, 0.05, coef, inf,
                       comms[r].param_peer_ptrs(r * shard))
        comms[r].params_barrier()
    for _ in range(2):
        _run_ranks(world, step)
    torch.cuda.synchronize()
    want = torch.cat([state[r][0] for r in range(world)]).to(torch.bfloat16)
    for r in range(world):
        assert torch.equal(comms[r].pbuf, want), f"rank {r}"
        assert comms[r].error_flag() == 0
