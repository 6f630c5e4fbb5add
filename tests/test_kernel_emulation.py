"""SIMT kernels executed on the CPU: the real ``csrc/*.cu`` kernel source is compiled for the host against a small CUDA
execution-model shim (``tests/emu/cuda_emu``: one std::thread per CUDA thread, barriers for __syncthreads / __syncwarp,
shuffles through a per-warp buffer) and compared with the fp32 PyTorch oracle.  This validates indexing, masking, the
online-softmax and the split / warp merges without a GPU; the hardware tests live in ``tests/test_ops_gpu.py``."""
import ctypes
import math
import os
import subprocess

import pytest
import torch

from megatron_llm_b200.ops.attention import attention_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


def _build(name, tmp_path_factory):
    out = os.path.join(str(tmp_path_factory.getbasetemp()), name + ".so")
    if not os.path.exists(out):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-shared", "-fPIC", "-pthread",
                               "-I" + os.path.join(EMU, "cuda_emu"), os.path.join(EMU, name + ".cpp"), "-o", out])
    return ctypes.CDLL(out)


@pytest.fixture(scope="module")
def decode_lib(tmp_path_factory):
    return _build("emu_attention_decode", tmp_path_factory)


def _strides(t):
    return (ctypes.c_longlong * 3)(t.stride(0), t.stride(1), t.stride(2))


def _decode(lib, q, k, v, window, splits, keys_per_split=None):
    b, sq, nq, hn = q.shape
    sk, nkv = k.size(1), k.size(2)
    if keys_per_split is None:
        keys_per_split = (((sk + splits - 1) // splits) + 31) // 32 * 32
        splits = (sk + keys_per_split - 1) // keys_per_split
    rows = b * nkv * splits * sq * (nq // nkv)
    part_o = torch.full((rows, hn), float("nan"))
    part_ml = torch.full((rows, 2), float("nan"))
    out = torch.full((b, sq, nq, hn), float("nan"), dtype=q.dtype)
    rc = lib.emu_attn_decode(
        ctypes.c_int(0 if q.dtype == torch.bfloat16 else 1), ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(k.data_ptr()),
        ctypes.c_void_p(v.data_ptr()), _strides(q), _strides(k), _strides(v), b, sq, sk, nq, nkv, hn,
        ctypes.c_int(-1 if window is None else window), ctypes.c_float(1.0 / math.sqrt(hn)), splits, keys_per_split,
        ctypes.c_void_p(part_o.data_ptr()), ctypes.c_void_p(part_ml.data_ptr()), ctypes.c_void_p(out.data_ptr()))
    assert rc == 0, rc
    return out


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("b,sq,sk,nq,nkv,hn,window,splits", [
    (2, 1, 77, 4, 4, 128, None, 2),       # MHA, ragged cache length, a half-empty last tile
    (1, 1, 300, 8, 2, 128, None, 3),      # GQA (4 rows per group)
    (2, 1, 200, 16, 1, 64, None, 1),      # MQA with 16 rows -> two passes of 8 rows, head_dim 64
    (1, 3, 130, 6, 2, 64, None, 2),       # a few query positions (bottom-right aligned causal mask), 9 rows per group
    (1, 1, 260, 4, 2, 128, 64, 4),        # sliding window: the first splits are fully masked
    (1, 2, 96, 2, 2, 128, 16, 1),         # window smaller than a tile
    (1, 5, 5, 2, 1, 64, None, 1),         # sq == sk (a short prompt)
])
def test_decode_attention_kernel_on_cpu_threads(decode_lib, dtype, b, sq, sk, nq, nkv, hn, window, splits):
    g = torch.Generator().manual_seed(1234 + sk)
    q = torch.randn(b, sq, nq, hn, generator=g).to(dtype)
    # k / v as views of a larger [s_max, b_max, nkv, hn] cache, like the inference path hands them over
    kmem = torch.randn(sk + 9, b + 1, nkv, hn, generator=g).to(dtype)
    vmem = torch.randn(sk + 9, b + 1, nkv, hn, generator=g).to(dtype)
    k = kmem[:sk, 1:1 + b].transpose(0, 1)
    v = vmem[:sk, 1:1 + b].transpose(0, 1)
    out = _decode(decode_lib, q, k, v, window, splits)
    ref = attention_reference(q.float(), k.float(), v.float(), causal=True, window=window)
    assert torch.isfinite(out.float()).all()
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert (out.float() - ref).abs().max().item() < tol


def test_decode_attention_kernel_empty_trailing_split(decode_lib):
    """More splits than the cache needs (keys_per_split * n_splits >> sk): the empty slices contribute nothing."""
    g = torch.Generator().manual_seed(7)
    q = torch.randn(1, 1, 2, 64, generator=g).bfloat16()
    k = torch.randn(1, 40, 2, 64, generator=g).bfloat16()
    v = torch.randn(1, 40, 2, 64, generator=g).bfloat16()
    out = _decode(decode_lib, q, k, v, None, splits=3, keys_per_split=32)
    ref = attention_reference(q.float(), k.float(), v.float(), causal=True)
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("p", [0.1, 0.5, 0.03])
def test_attention_dropout_mask_replica_matches_kernel_source(tmp_path_factory, p):
    """The Python replica of the dropout mask (used as the oracle of the GPU dropout tests) equals the C++ the kernels
    compile (csrc/attention_dropout.cuh), and the mask has the requested rate without row / column structure."""
    from megatron_llm_b200.ops.attention import dropout_keep_mask, dropout_threshold
    lib = _build("emu_attention_dropout", tmp_path_factory)
    seed, n_bh, rows, keys = 0x1234_5678_9ABC_DEF1, 6, 300, 260
    keep = torch.zeros(n_bh, rows, keys, dtype=torch.uint8)
    inv = ctypes.c_float()
    thr = lib.emu_dropout_keep(ctypes.c_float(p), ctypes.c_ulonglong(seed), n_bh, rows, keys,
                               ctypes.c_void_p(keep.data_ptr()), ctypes.byref(inv))
    t, inv_keep = dropout_threshold(p)
    assert thr == t and abs(inv.value - inv_keep) < 1e-6
    mine = dropout_keep_mask(seed, p, 2, 3, rows, keys)
    assert torch.equal(mine.view(n_bh, rows, keys), keep.bool())
    rate = 1.0 - keep.float().mean().item()
    assert abs(rate - t / 256.0) < 0.01
    # no structure: every row and every column drops at about the same rate, different heads / seeds differ
    assert (1.0 - keep.float().mean(dim=(0, 2)) - t / 256.0).abs().max().item() < 0.06
    assert (1.0 - keep.float().mean(dim=(0, 1)) - t / 256.0).abs().max().item() < 0.06
    assert not torch.equal(keep[0], keep[1])
    other = dropout_keep_mask(seed + 1, p, 2, 3, rows, keys)
    assert (other != mine).float().mean().item() > 0.5 * min(t, 256 - t) / 256.0


def test_attention_dropout_backward_formulas():
    """The tile math the tcgen05 backward kernels implement for dropout (csrc/attention_bwd_sm100.cu):
    O = (P o Z) V with Z = keep / (1 - p);  delta = rowsum(dO o O);  dV = (P o Z)^T dO;  dS = P o (Z o (dO V^T) - delta);
    dQ = scale dS K;  dK = scale dS^T Q -- checked against autograd through the fp32 oracle with the same mask."""
    from megatron_llm_b200.ops.attention import dropout_keep_mask, dropout_threshold
    torch.manual_seed(3)
    b, s, n, hn, p, seed = 2, 48, 3, 16, 0.25, 0xABCDEF0123456789 & 0x7FFFFFFFFFFFFFFF
    q, k, v, do = (torch.randn(b, s, n, hn, dtype=torch.float64) for _ in range(4))
    keep = dropout_keep_mask(seed, p, b, n, s, s)
    qa, ka, va = (t.clone().requires_grad_() for t in (q, k, v))
    out = attention_reference(qa, ka, va, True, None, None, p, keep)          # float() inside: fp32 oracle
    out.backward(do.float())
    scale = 1.0 / math.sqrt(hn)
    Q, K, V, dO = (t.permute(0, 2, 1, 3) for t in (q, k, v, do))                # [b, n, s, hn]
    S = (Q @ K.transpose(-1, -2)) * scale
    causal = torch.ones(s, s, dtype=torch.bool).tril()
    P = torch.softmax(S.masked_fill(~causal, float("-inf")), dim=-1)
    Z = keep.double() * dropout_threshold(p)[1]
    O = (P * Z) @ V
    delta = (dO * O).sum(-1, keepdim=True)
    dV = (P * Z).transpose(-1, -2) @ dO
    dS = P * (Z * (dO @ V.transpose(-1, -2)) - delta)
    dQ = scale * dS @ K
    dK = scale * dS.transpose(-1, -2) @ Q
    for mine, ref in ((O, out.permute(0, 2, 1, 3)), (dQ, qa.grad.permute(0, 2, 1, 3)), (dK, ka.grad.permute(0, 2, 1, 3)),
                      (dV, va.grad.permute(0, 2, 1, 3))):
        assert (mine - ref.double()).abs().max().item() < 2e-5
