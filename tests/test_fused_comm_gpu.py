"""Fused GEMM+collective kernels and the peer-memory DP reduction vs. the unfused NCCL path (needs >= 2 B200s).
The same kernels are exercised on ONE GPU by tests/test_fused_loopback_gpu.py (virtual ranks on one device)."""
import os

import pytest
import torch
import torch.distributed as dist

from tests.dist_utils import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _tp_kernels(rank, world):
    from megatron_llm_b200.parallel import state as ps
    from megatron_llm_b200.parallel.symm import TPCommunicator
    ps.initialize_model_parallel(world, 1)
    dev = torch.device("cuda", rank)
    torch.manual_seed(100 + rank)
    K, N = 512, 768
    comm = TPCommunicator(ps.get_tensor_model_parallel_group(), max_rows_per_rank=512, max_k=2048, max_n=2048,
                          num_comm_ctas=4, all_reduce_n=1024, nvls_ag_k=2048)   # (NVLS gather only with MLB200_AG_NVLS=1)
    group = ps.get_tensor_model_parallel_group()
    # m = 256 takes the 2-CTA (cta_group::2) kernels, m = 128 the 1-CTA ones; alternating them on one communicator
    # also checks that the arrival / epoch accounting is shared correctly between the two variants
    for it, m in enumerate((256, 256, 128, 256, 128, 128)):
        # ---- all-gather -> GEMM (W [N,K]) and the transposed-weight form (W [K,N])
        x = torch.randn(m, K, device=dev, dtype=torch.bfloat16)
        torch.manual_seed(7 + it)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
        wt = torch.randn(K, N, device=dev, dtype=torch.bfloat16) * 0.05
        full = torch.empty(world * m, K, device=dev, dtype=torch.bfloat16)
        dist.all_gather_into_tensor(full, x, group=group)
        out, gathered = comm.ag_gemm(x, w, False)
        torch.cuda.synchronize()
        assert torch.equal(gathered, full), f"gathered mismatch it={it}"
        ref = full.float() @ w.float().t()
        assert (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max(), f"ag_gemm it={it}"
        out2, _ = comm.ag_gemm(x, wt, True)
        ref2 = full.float() @ wt.float()
        assert (out2.float() - ref2).abs().max() <= 2e-2 * ref2.abs().max(), f"ag_gemm(T) it={it}"
        # ---- GEMM -> reduce-scatter
        M = world * m
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        part = (a.float() @ w.float().t())
        refs = torch.empty(m, N, device=dev, dtype=torch.float32)
        dist.reduce_scatter_tensor(refs, part, group=group)
        got = comm.gemm_rs(a, w, False)
        torch.cuda.synchronize()
        assert (got.float() - refs).abs().max() <= 3e-2 * refs.abs().max(), f"gemm_rs it={it}"
        part2 = (a.float() @ wt.float())
        dist.reduce_scatter_tensor(refs, part2, group=group)
        got2 = comm.gemm_rs(a, wt, True)
        assert (got2.float() - refs).abs().max() <= 3e-2 * refs.abs().max(), f"gemm_rs(T) it={it}"
        # ---- GEMM -> all-reduce (non-sequence-parallel form)
        full_sum = part.clone()
        dist.all_reduce(full_sum, group=group)
        got3 = comm.gemm_ar(a, w, False)
        torch.cuda.synchronize()
        assert got3.shape == (M, N)
        assert (got3.float() - full_sum).abs().max() <= 3e-2 * full_sum.abs().max(), f"gemm_ar it={it}"
    assert comm.error_flag() == 0, "a spin-wait timed out"
    ps.destroy_model_parallel()


def test_fused_tp_kernels_match_nccl():
    run_distributed(_tp_kernels, 2, backend="nccl")


def _tp_kernels_nvls(rank, world):
    os.environ["MLB200_AG_NVLS"] = "1"       # read when the communicator is built
    _tp_kernels(rank, world)


def test_fused_tp_kernels_nvls_all_gather_match_nccl():
    """Same checks with the NVLS transport of the all-gather (pusher CTAs multimem.st the shard into every rank's gather
    buffer); m = 128 calls still take the pull kernel, so both transports alternate on one communicator.  (Falls back
    to the pull kernel, i.e. repeats the test above, on a box without multicast support.)"""
    run_distributed(_tp_kernels_nvls, min(torch.cuda.device_count(), 4), backend="nccl")


def _tp_kernels_nvls_in_graph(rank, world):
    os.environ["MLB200_AG_NVLS"] = "1"
    _tp_kernels_in_graph(rank, world)


def test_fused_tp_kernels_nvls_replay_in_cuda_graph():
    run_distributed(_tp_kernels_nvls_in_graph, 2, backend="nccl")


def _tp_kernels_with_skew(rank, world):
    """Shake the flag protocol: every rank delays its launches by a different, changing amount (device-side spin on the
    launching stream), so the READY / ACK / ARRIVED / FREE handshakes see peers that are early, late, or a whole call
    behind.  Results must still match NCCL and no bounded spin may time out."""
    import random
    from megatron_llm_b200.parallel import state as ps
    from megatron_llm_b200.parallel.symm import TPCommunicator
    ps.initialize_model_parallel(world, 1)
    dev = torch.device("cuda", rank)
    group = ps.get_tensor_model_parallel_group()
    m, K, N = 256, 512, 768
    comm = TPCommunicator(group, max_rows_per_rank=512, max_k=2048, max_n=2048, num_comm_ctas=4)
    rnd = random.Random(1234 + rank)
    torch.manual_seed(3)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
    results = []
    for it in range(12):
        torch.manual_seed(100 * it + rank)
        x = torch.randn(m, K, device=dev, dtype=torch.bfloat16)
        a = torch.randn(world * m, K, device=dev, dtype=torch.bfloat16)
        torch.cuda._sleep(rnd.randrange(0, 3_000_000))          # up to ~1.5 ms of skew before the all-gather GEMM
        out, gathered = comm.ag_gemm(x, w, False)
        torch.cuda._sleep(rnd.randrange(0, 3_000_000))
        got = comm.gemm_rs(a, w, False)
        results.append((x, a, out, gathered, got))               # no host sync inside the loop
    torch.cuda.synchronize()
    for it, (x, a, out, gathered, got) in enumerate(results):
        full = torch.empty(world * m, K, device=dev, dtype=torch.bfloat16)
        dist.all_gather_into_tensor(full, x, group=group)
        assert torch.equal(gathered, full), f"it {it}: gathered mismatch"
        ref = full.float() @ w.float().t()
        assert (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max(), f"it {it}: ag_gemm"
        refs = torch.empty(m, N, device=dev, dtype=torch.float32)
        dist.reduce_scatter_tensor(refs, a.float() @ w.float().t(), group=group)
        assert (got.float() - refs).abs().max() <= 3e-2 * refs.abs().max(), f"it {it}: gemm_rs"
    assert comm.error_flag() == 0, "a spin-wait timed out"
    ps.destroy_model_parallel()


def test_fused_tp_kernels_tolerate_rank_skew():
    run_distributed(_tp_kernels_with_skew, 2, backend="nccl")


def _tp_kernels_in_graph(rank, world):
    """The fused kernels captured in a CUDA graph: every replay must continue the live epoch sequence, also when eager
    calls run in between (which flips the receive-slot parity the captured reduce-scatter was recorded with)."""
    from megatron_llm_b200.parallel import state as ps
    from megatron_llm_b200.parallel.symm import TPCommunicator
    ps.initialize_model_parallel(world, 1)
    dev = torch.device("cuda", rank)
    group = ps.get_tensor_model_parallel_group()
    m, K, N = 256, 512, 768
    comm = TPCommunicator(group, max_rows_per_rank=512, max_k=2048, max_n=2048, num_comm_ctas=4, nvls_ag_k=2048)
    torch.manual_seed(5)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
    x = torch.randn(m, K, device=dev, dtype=torch.bfloat16)
    a = torch.randn(world * m, K, device=dev, dtype=torch.bfloat16)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for _ in range(3):                      # eager calls first: the capture starts from non-zero counters
            comm.ag_gemm(x, w, False)
            comm.gemm_rs(a, w, False)
        torch.cuda.synchronize()
        before = comm.counters()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
            out, gathered = comm.ag_gemm(x, w, False)
            got = comm.gemm_rs(a, w, False)     # one reduce-scatter per replay: odd epoch advance
        advance = comm.end_capture(before)
        assert advance[0] == 1 and advance[1] == 1 and advance[2] > 0
        for it in range(6):
            torch.manual_seed(1000 * it + rank)
            x.copy_(torch.randn(m, K, device=dev, dtype=torch.bfloat16))
            a.copy_(torch.randn(world * m, K, device=dev, dtype=torch.bfloat16))
            comm.begin_replay(before, advance)
            graph.replay()
            torch.cuda.synchronize()
            full = torch.empty(world * m, K, device=dev, dtype=torch.bfloat16)
            dist.all_gather_into_tensor(full, x, group=group)
            assert torch.equal(gathered, full), f"replay {it}: gathered mismatch"
            ref = full.float() @ w.float().t()
            assert (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max(), f"replay {it}: ag_gemm"
            refs = torch.empty(m, N, device=dev, dtype=torch.float32)
            dist.reduce_scatter_tensor(refs, a.float() @ w.float().t(), group=group)
            assert (got.float() - refs).abs().max() <= 3e-2 * refs.abs().max(), f"replay {it}: gemm_rs"
            if it % 3 == 1:                     # an eager call between two replays
                e = comm.gemm_rs(a, w, False)
                torch.cuda.synchronize()
                assert (e.float() - refs).abs().max() <= 3e-2 * refs.abs().max(), f"eager after replay {it}"
    assert comm.error_flag() == 0, "a spin-wait timed out"
    ps.destroy_model_parallel()


def test_fused_tp_kernels_replay_in_cuda_graph():
    run_distributed(_tp_kernels_in_graph, 2, backend="nccl")


def _dp_reduce(rank, world, nvls):
    os.environ["MLB200_DP_NVLS"] = "1" if nvls else "0"
    from megatron_llm_b200.parallel import state as ps
    from megatron_llm_b200.parallel.symm import DPCommunicator
    ps.initialize_model_parallel(1, 1)
    dev = torch.device("cuda", rank)
    n = 1 << 20
    comm = DPCommunicator(ps.get_data_parallel_group(), n)
    if nvls and not comm.use_nvls:
        ps.destroy_model_parallel()
        return                               # no multicast mapping on this box: nothing to test
    half = n // 2
    for it in range(4):
        reduce_scatter = it % 2 == 1
        torch.manual_seed(rank * 10 + it)
        g = torch.randn(n, device=dev)
        ref = g.clone()
        dist.all_reduce(ref, group=ps.get_data_parallel_group())
        ref /= world
        comm.buffer.copy_(g)
        h = comm.reduce_bucket(comm.buffer[:half], 0, n, reduce_scatter=reduce_scatter)
        h2 = comm.reduce_bucket(comm.buffer[half:], half, n, reduce_scatter=reduce_scatter)
        h.wait(); h2.wait()
        torch.cuda.synchronize()
        if reduce_scatter:                   # rank r owns slice r of each bucket
            sl = half // world
            for base in (0, half):
                a, b = base + rank * sl, base + (rank + 1) * sl
                assert torch.allclose(comm.buffer[a:b], ref[a:b], atol=1e-5), f"dp reduce-scatter it={it}"
        else:
            assert torch.allclose(comm.buffer, ref, atol=1e-5), f"dp all-reduce it={it}"
        dist.barrier()
    assert comm.error_flag() == 0
    ps.destroy_model_parallel()


@pytest.mark.parametrize("nvls", [False, True], ids=["peer_pointers", "nvls_multimem"])
def test_dp_peer_memory_reduction(nvls):
    """Bucket all-reduce / reduce-scatter fused with the 1/DP scale: peer-pointer kernel and the NVLS kernel
    (multimem.ld_reduce in-switch sum + multimem.st write-back)."""
    run_distributed(_dp_reduce, min(torch.cuda.device_count(), 4), nvls, backend="nccl")
