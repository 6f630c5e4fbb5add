"""The tcgen05 GEMM kernels (plain mode) executed on the CPU on the functional model of TMA / mbarrier / tensor memory /
tcgen05.mma (tests/emu/cuda_emu/tcgen05_model.h, see tests/test_attention_kernel_model.py):
  * the 1-CTA kernel (csrc/gemm_sm100.cuh): NT / NN / TN operand layouts (K-major and MN-major shared-memory
    descriptors for A and B), the four epilogues, both tile widths, ragged M / N / K edges (TMA zero fill, guarded
    stores), fp16 operands, a persistent CTA walking several tiles;
  * the 2-CTA kernel (csrc/gemm2_sm100.cu, ``cta_group::2``, the GEMM of the training step): the two CTAs of a cluster
    run concurrently -- each loads its half of A and B, the bytes of both complete on the leader's barriers, the leader's
    MMAs read both shared memories and write both tensor memories, commits are multicast, the epilogue warps arrive
    remotely, results leave through swizzled shared memory and TMA stores / fp32 reduce-adds clipped by the tensor map;
and their barrier protocols under ThreadSanitizer.  The fused all-gather / reduce-scatter modes need co-resident puller
CTAs and peers and are not run here."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import host_build  # noqa: E402

EPI_16, EPI_F32_ACCUM, EPI_F32, EPI_16_ACCUM = 0, 1, 2, 3


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    return ctypes.CDLL(host_build.build(["gemm_sm100.cu", "gemm2_sm100.cu"], str(tmp_path_factory.mktemp("emu_gemm"))))


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,epi,block_n,dtype", [
    (128, 128, 64, 0, 0, EPI_16, 128, torch.bfloat16),          # NT (forward of a Linear), one tile
    (200, 264, 136, 0, 0, EPI_16, 128, torch.bfloat16),         # ragged M / N / K
    (256, 256, 128, 0, 1, EPI_16, 256, torch.bfloat16),         # NN (dgrad): B is MN-major, 256-wide tiles
    (136, 128, 256, 1, 1, EPI_F32, 128, torch.bfloat16),        # TN (wgrad): A and B MN-major, fp32 output
    (136, 200, 128, 1, 1, EPI_F32_ACCUM, 128, torch.bfloat16),  # ... accumulated into an fp32 main_grad
    (128, 136, 72, 0, 0, EPI_16_ACCUM, 128, torch.bfloat16),    # 16-bit accumulate epilogue
    (128, 256, 64, 0, 0, EPI_16, 256, torch.float16),           # fp16 operands and output
    (640, 128, 192, 0, 0, EPI_16, 128, torch.bfloat16),         # persistent CTAs: 5 tiles on 4 "SMs"
])
def test_gemm_kernel_on_the_functional_model(lib, M, N, K, a_mn, b_mn, epi, block_n, dtype):
    torch.manual_seed(M + N + K)
    A, B = torch.randn(M, K).to(dtype), torch.randn(N, K).to(dtype)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    out_dtype = dtype if epi in (EPI_16, EPI_16_ACCUM) else torch.float32
    before = torch.randn(M, N).to(out_dtype) if epi in (EPI_F32_ACCUM, EPI_16_ACCUM) else None
    c = before.clone() if before is not None else torch.full((M, N), float("nan"), dtype=out_dtype)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.mlb_gemm_bf16(p(a), p(b), p(c), M, N, K, a.stride(0), b.stride(0), c.stride(0), a_mn, b_mn, epi, block_n,
                             int(dtype == torch.float16), 4, None) == 0
    ref = A.float() @ B.float().t()
    if before is not None:
        ref = ref + before.float()
    err = ((c.float() - ref).norm() / ref.norm()).item()
    assert err < (1e-5 if out_dtype == torch.float32 else (3e-3 if dtype == torch.bfloat16 else 5e-4)), err


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,epi,dtype", [
    (256, 256, 64, 0, 0, EPI_16, torch.bfloat16),               # one 256 x 256 tile on one CTA pair
    (520, 264, 200, 0, 1, EPI_16, torch.bfloat16),              # NN, ragged: clipped TMA stores, zero-filled loads
    (256, 512, 192, 1, 1, EPI_F32, torch.bfloat16),             # TN, fp32 output boxes
    (264, 256, 128, 1, 1, EPI_F32_ACCUM, torch.bfloat16),       # wgrad: fp32 reduce-add into main_grad
    (256, 264, 72, 0, 0, EPI_16_ACCUM, torch.bfloat16),         # 16-bit accumulate (register epilogue)
    (768, 512, 128, 0, 0, EPI_16, torch.float16),               # 6 tiles on 2 pairs, fp16
])
def test_two_cta_gemm_kernel_on_the_functional_model(lib, M, N, K, a_mn, b_mn, epi, dtype):
    torch.manual_seed(M + N + K)
    A, B = torch.randn(M, K).to(dtype), torch.randn(N, K).to(dtype)
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    out_dtype = dtype if epi in (EPI_16, EPI_16_ACCUM) else torch.float32
    before = torch.randn(M, N).to(out_dtype) if epi in (EPI_F32_ACCUM, EPI_16_ACCUM) else None
    c = before.clone() if before is not None else torch.full((M, N), float("nan"), dtype=out_dtype)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.mlb_gemm_bf16_2cta(p(a), p(b), p(c), M, N, K, a.stride(0), b.stride(0), c.stride(0), a_mn, b_mn, epi,
                                  int(dtype == torch.float16), 4, None) == 0
    ref = A.float() @ B.float().t()
    if before is not None:
        ref = ref + before.float()
    err = ((c.float() - ref).norm() / ref.norm()).item()
    assert err < (1e-5 if out_dtype == torch.float32 else (3e-3 if dtype == torch.bfloat16 else 5e-4)), err


def test_gemm_barrier_protocols_under_thread_sanitizer(tmp_path):
    exe = host_build.build_race_driver(["gemm_sm100.cu"], str(tmp_path / "one"), name="gemm_race", driver="gemm_race_driver.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
    exe = host_build.build_race_driver(["gemm_sm100.cu", "gemm2_sm100.cu"], str(tmp_path / "two"), name="gemm2_race",
                                       driver="gemm2_race_driver.cpp")
    for epi in ("0", "1", "3"):
        r = subprocess.run([exe, epi], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (epi, r.stderr[-3000:])


@pytest.mark.parametrize("seed", [1, 2])
def test_tcgen05_kernels_survive_schedule_fuzzing(tmp_path_factory, seed):
    """Attention (both forward kernels + backward, window, dropout) and both GEMM kernels with random stalls in front of
    every barrier operation, TMA load, MMA and tensor-memory access (MLB_EMU_CHAOS): warps drift apart by whole tiles,
    so a protocol that relied on relative speed -- like the o_done wait of the single-tile attention kernel before its
    fix, which fails this check every time -- produces wrong numbers or hangs."""
    import json
    base = tmp_path_factory.getbasetemp()
    attn = os.path.join(base, "chaos_attn", "emu_kernels.so")
    gm = os.path.join(base, "chaos_gemm", "emu_kernels.so")
    if not os.path.exists(attn):
        host_build.build(["attention_sm100.cu", "attention_bwd_sm100.cu"], os.path.dirname(attn))
        host_build.build(["gemm_sm100.cu", "gemm2_sm100.cu"], os.path.dirname(gm))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "chaos_check.py"), attn, gm],
                       env=dict(os.environ, MLB_EMU_CHAOS=str(seed)), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert all(v == v and v < 6e-3 for v in res.values()), res


# ------------------------------------------------------------------------------------------------------------------
# fused GEMM -> reduce-scatter / all-reduce between emulated tensor-parallel ranks (tests/emu/fused_rank.cpp)
# ------------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def fused_rank_exe(tmp_path_factory):
    return host_build.build_executable(["gemm_sm100.cu", "gemm2_sm100.cu"], "fused_rank.cpp",
                                       str(tmp_path_factory.mktemp("emu_fused")), "fused_rank")


def _fused_ranks(exe, d, world, m, N, K, mode, calls, chaos=None, ctas=1, graph_state=False):
    import numpy as np
    os.makedirs(d, exist_ok=True)
    for r in range(world):
        np.zeros(2 * world * m * N, dtype=np.uint16).tofile(os.path.join(d, f"slots{r}.bin"))
        np.zeros(2 * world * m * N, dtype=np.uint16).tofile(os.path.join(d, f"arout{r}.bin"))
        np.zeros(64, dtype=np.int32).tofile(os.path.join(d, f"pad{r}.bin"))
    env = dict(os.environ, **({"MLB_EMU_CHAOS": str(chaos)} if chaos else {}),
               **({"MLB_EMU_CONCURRENT_BLOCKS": "1"} if ctas > 1 else {}),
               **({"FUSED_RANK_GRAPH_STATE": "1"} if graph_state else {}))
    procs = [subprocess.Popen([exe, str(d), str(r), str(world), str(m), str(N), str(K), mode, str(calls), str(ctas)],
                              env=env, stderr=subprocess.PIPE, text=True) for r in range(world)]
    return [(p.wait(timeout=900), p.stderr.read()[-300:]) for p in procs]


@pytest.mark.parametrize("mode,world,m,N,K,calls,chaos,ctas", [
    ("rs1", 2, 256, 128, 128, 3, None, 1),   # 1-CTA kernel: epilogue stores into the peer's slot, arrival counters,
    ("rs1", 3, 128, 128, 64, 2, 5, 1),       #   slot reduction, PAD_RS_FREE on the third call; with schedule fuzzing
    ("ar1", 3, 128, 128, 64, 2, None, 1),    # GEMM -> all-reduce: the reduced slice stored into every rank's output
    ("rs2", 2, 256, 256, 128, 3, None, 1),   # 2-CTA kernel: TMA stores into peer slots, signalling one tile late
    ("rs2", 2, 512, 256, 64, 2, 9, 1),
    ("rs1", 2, 256, 128, 128, 3, None, 3),   # several CTAs (pairs) per rank running concurrently: the slot reduction is
    ("rs2", 2, 512, 256, 128, 3, 6, 2),      #   spread over them and the last one out hands the slots back
    ("ar1", 3, 256, 128, 64, 2, 2, 2),
])
def test_fused_gemm_reduce_scatter_between_emulated_ranks(fused_rank_exe, tmp_path, mode, world, m, N, K, calls, chaos,
                                                          ctas):
    """Row-parallel forward Y = sum_r X_r W_r^T with the reduce-scatter (or all-reduce) fused into the GEMM kernel: every
    rank is a process running the real kernel source on the functional model (one CTA / one CTA pair per rank), receive
    slots and signal pads are files all ranks map.  Each rank checks its rows against a reference it recomputes from
    the seeds and that none of its bounded spins timed out; several calls alternate the slot parities."""
    res = _fused_ranks(fused_rank_exe, tmp_path, world, m, N, K, mode, calls, chaos, ctas)
    assert all(rc == 0 for rc, _ in res), res


def _ag_ranks(exe, d, world, m, N, K, pullers, calls, chaos=None, two_cta=False):
    """``two_cta``: False (1-CTA kernel), True (2-CTA kernel, pull transport) or "nvls" (2-CTA kernel, push transport)."""
    import numpy as np
    os.makedirs(d, exist_ok=True)
    for r in range(world):
        np.zeros(m * K, dtype=np.uint16).tofile(os.path.join(d, f"shard{r}.bin"))
        np.zeros(64, dtype=np.int32).tofile(os.path.join(d, f"pad{r}.bin"))
        if two_cta == "nvls":
            np.zeros(2 * world * m * K, dtype=np.uint16).tofile(os.path.join(d, f"gather{r}.bin"))
            np.zeros(2 * (world * m // 128), dtype=np.int32).tofile(os.path.join(d, f"gflags{r}.bin"))
    env = dict(os.environ, MLB_EMU_CONCURRENT_BLOCKS="1", **({"MLB_EMU_CHAOS": str(chaos)} if chaos else {}))
    procs = [subprocess.Popen([exe, str(d), str(r), str(world), str(m), str(N), str(K), str(pullers), str(calls)] +
                              ([two_cta if two_cta == "nvls" else "2cta"] if two_cta else []), env=env,
                              stderr=subprocess.PIPE, text=True)
             for r in range(world)]
    return [(p.wait(timeout=900), p.stderr.read()[-300:]) for p in procs]


@pytest.mark.parametrize("world,m,N,K,pullers,calls,chaos,two_cta", [
    (2, 256, 256, 128, 2, 3, None, False), (3, 256, 128, 320, 3, 2, 4, False), (2, 128, 136, 72, 1, 2, 8, False),
    (2, 256, 256, 128, 2, 2, None, True),          # 2-CTA kernel: puller clusters next to compute clusters
    (2, 512, 264, 72, 2, 3, 3, True),
    (2, 256, 256, 128, 2, 4, None, "nvls"),        # push transport: multimem.st into every rank's gather buffer
    (2, 512, 256, 64, 4, 3, 5, "nvls"),
])
def test_fused_all_gather_gemm_between_emulated_ranks(tmp_path_factory, tmp_path, world, m, N, K, pullers, calls, chaos,
                                                      two_cta):
    """Column-parallel forward out_r = all_gather(X) W_r^T with the all-gather fused into the GEMM launch: puller CTAs
    (bulk copies of the peers' published shards through shared memory, one flag per 128-row chunk, read
    acknowledgements) and compute CTAs (TMA producers that wait for the chunk they are about to read) run
    concurrently inside every rank; ranks are processes over mapped files."""
    base = str(tmp_path_factory.getbasetemp())
    exe = os.path.join(base, "emu_fused_ag", "fused_ag_rank")
    if not os.path.exists(exe):
        host_build.build_executable(["gemm_sm100.cu", "gemm2_sm100.cu"], "fused_ag_rank.cpp", os.path.dirname(exe), "fused_ag_rank")
    res = _ag_ranks(exe, tmp_path, world, m, N, K, pullers, calls, chaos, two_cta)
    assert all(rc == 0 for rc, _ in res), res


def test_fused_kernels_between_ranks_under_thread_sanitizer(tmp_path):
    """The same rank programs built with ThreadSanitizer (it sees the accesses inside a rank: pullers vs. compute CTAs,
    epilogue vs. slot reduction; the peers' stores arrive through the mapped files)."""
    files = ["gemm_sm100.cu", "gemm2_sm100.cu"]
    ag = host_build.build_race_driver(files, str(tmp_path / "ag"), name="ag_rank_tsan", driver="fused_ag_rank.cpp")
    res = _ag_ranks(ag, tmp_path / "ag_run", 2, 128, 128, 128, 2, 2)
    assert all(rc == 0 and "ThreadSanitizer" not in err for rc, err in res), res
    rs = host_build.build_race_driver(files, str(tmp_path / "rs"), name="rs_rank_tsan", driver="fused_rank.cpp")
    for mode, m, N in (("rs1", 128, 128), ("rs2", 256, 256)):
        res = _fused_ranks(rs, tmp_path / ("rs_run_" + mode), 2, m, N, 64, mode, 2)
        assert all(rc == 0 and "ThreadSanitizer" not in err for rc, err in res), (mode, res)


def test_fused_kernel_with_a_lost_peer_times_out_instead_of_hanging(fused_rank_exe, tmp_path):
    """A tensor-parallel peer that dies before delivering its tiles: the bounded spins of the fused GEMM -> reduce-scatter
    kernel end, the timeout is recorded in the signal pad (``symm.check_timeouts`` makes it fatal once per training step)
    and the kernel retires -- the rank program reports it (exit code 4) instead of hanging."""
    import numpy as np
    world, m, N, K = 2, 128, 128, 64
    for r in range(world):
        np.zeros(2 * world * m * N, dtype=np.uint16).tofile(os.path.join(tmp_path, f"slots{r}.bin"))
        np.zeros(2 * world * m * N, dtype=np.uint16).tofile(os.path.join(tmp_path, f"arout{r}.bin"))
        np.zeros(64, dtype=np.int32).tofile(os.path.join(tmp_path, f"pad{r}.bin"))
    open(os.path.join(tmp_path, "ready1"), "w").close()            # rank 1 "arrived" at the rendezvous, then vanished
    r = subprocess.run([fused_rank_exe, str(tmp_path), "0", str(world), str(m), str(N), str(K), "rs1", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 4 and "timed out" in r.stderr, (r.returncode, r.stderr[-500:])
    assert np.fromfile(os.path.join(tmp_path, "pad0.bin"), dtype=np.int32)[32] == 1


@pytest.mark.parametrize("mode,m,N", [("rs1", 128, 128), ("rs2", 256, 256)])
def test_fused_kernels_follow_device_resident_epoch_offsets(fused_rank_exe, tmp_path, mode, m, N):
    """What a replayed CUDA graph does (``GraphedMicrobatch``): the kernel arguments stay those of the capture and the
    live epoch / cumulative arrival count come from ``GemmComm::state`` in device memory -- five calls over both
    parities with frozen arguments."""
    res = _fused_ranks(fused_rank_exe, tmp_path, 2, m, N, 64, mode, 5, graph_state=True)
    assert all(rc == 0 for rc, _ in res), res
