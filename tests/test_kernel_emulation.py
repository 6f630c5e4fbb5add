"""SIMT kernels executed on the CPU: the real ``csrc/*.cu`` kernel source is compiled for the host against a small CUDA
execution-model shim (``tests/emu/cuda_emu``: one std::thread per CUDA thread, barriers for __syncthreads / __syncwarp,
shuffles through a per-warp buffer; ``tests/emu/host_build.py`` only rewrites the ``<<<...>>>`` launch syntax, so the
real ``extern "C"`` launchers run too) and compared with the fp32 PyTorch oracle.  This validates indexing, masking,
reductions, the online softmax and the split / warp merges without a GPU; the hardware tests live in
``tests/test_ops_gpu.py``.  (The tcgen05 / TMA kernels cannot be run this way.)"""
import ctypes
import math
import os
import sys

import pytest
import torch

from megatron_llm_b200.ops.attention import attention_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
sys.path.insert(0, EMU)
import host_build  # noqa: E402

DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}       # csrc/common.cuh DType


def _ptr(t):
    return ctypes.c_void_p(None if t is None else t.data_ptr())


@pytest.fixture(scope="module")
def kernels(tmp_path_factory):
    """The SIMT translation units (kernels AND their extern "C" launchers) compiled for the host."""
    so = host_build.build(["attention_decode.cu", "ce.cu", "softmax.cu", "norm.cu", "elementwise.cu"],
                          str(tmp_path_factory.mktemp("emu")))
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def decode_lib(kernels):
    return kernels


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
    rc = lib.mlb_attn_decode(
        ctypes.c_int(DT[q.dtype]), _ptr(q), _ptr(k), _ptr(v), _strides(q), _strides(k), _strides(v), b, sq, sk, nq, nkv,
        hn, ctypes.c_int(-1 if window is None else window), ctypes.c_float(1.0 / math.sqrt(hn)), splits,
        keys_per_split, _ptr(part_o), _ptr(part_ml), _ptr(out), None)
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
    out = str(tmp_path_factory.mktemp("emu_dropout"))
    lib = ctypes.CDLL(host_build.build([], out, "emu_dropout", [os.path.join(EMU, "emu_attention_dropout.cpp")]))
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


# ------------------------------------------------------------------------------------------------------------------
# the other SIMT kernels, through their real launchers (grid / block selection and dtype dispatch included)
# ------------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dtype,V", [(torch.bfloat16, 1000), (torch.float32, 517 + 3)])
def test_cross_entropy_kernels_on_cpu_threads(kernels, dtype, V):
    """csrc/ce.cu: pass 1 statistics (max, sum-exp, target logit, sum of logits) and the in-place gradient pass, with
    a vocabulary shard offset, a target outside the shard and label smoothing."""
    torch.manual_seed(0)
    rows, start, smoothing, vocab = 5, 200, 0.1, 4096
    stride = (V + 7) // 8 * 8
    buf = torch.zeros(rows, stride, dtype=dtype)
    buf[:, :V] = (torch.randn(rows, V) * 3).to(dtype)
    target = torch.tensor([start + 3, start + V - 1, 5, start, start + 17])      # 5 is owned by another shard
    stats = torch.zeros(rows, 4)
    assert kernels.mlb_ce_stats(DT[dtype], _ptr(buf), _ptr(target), _ptr(stats), rows, V, start,
                                ctypes.c_longlong(stride), None) == 0
    x = buf[:, :V].float()
    m = x.max(-1).values
    assert torch.allclose(stats[:, 0], m)
    assert torch.allclose(stats[:, 1], (x - m[:, None]).exp().sum(-1), rtol=1e-4)
    local = target - start
    owned = (local >= 0) & (local < V)
    tl = torch.where(owned, x.gather(1, local.clamp(0, V - 1)[:, None])[:, 0], torch.zeros(rows))
    assert torch.allclose(stats[:, 2], tl)
    assert torch.allclose(stats[:, 3], x.sum(-1), rtol=1e-4, atol=1e-3)
    # backward with this shard standing for the whole softmax: grad = (softmax - (1 - sm) onehot - sm / vocab) * g
    M, logS, g = stats[:, 0].contiguous(), stats[:, 1].log().contiguous(), torch.rand(rows) + 0.5
    out = torch.empty_like(buf)
    assert kernels.mlb_ce_bwd(DT[dtype], _ptr(buf), _ptr(out), _ptr(target), _ptr(M), _ptr(logS), _ptr(g), rows, V,
                              start, ctypes.c_float(smoothing), vocab, ctypes.c_longlong(stride), None) == 0
    ref = torch.softmax(x, -1) - smoothing / vocab
    ref[owned, local[owned]] -= 1.0 - smoothing
    ref = ref * g[:, None]
    assert (out[:, :V].float() - ref).abs().max().item() < (2e-2 if dtype == torch.bfloat16 else 1e-5)


@pytest.mark.parametrize("mode,sq,sk", [(0, 5, 40), (2, 24, 24), (2, 3, 70), (1, 6, 300), (2, 4, 1100)])
def test_softmax_kernels_on_cpu_threads(kernels, mode, sq, sk):
    """csrc/softmax.cu: plain / padding-mask / causal (bottom-right aligned) scaled softmax and the in-place backward,
    across the three block sizes the launcher picks."""
    torch.manual_seed(1)
    b, n, scale = 2, 2, 0.37
    x = torch.randn(b, n, sq, sk)
    mask = None
    if mode == 1:
        mask = (torch.rand(b, 1, sq, sk) < 0.3).to(torch.uint8)
        mask[0, 0, 1] = 1                                                  # a fully masked row -> zeros
    y = torch.empty_like(x)
    assert kernels.mlb_softmax_fwd(DT[x.dtype], _ptr(x), _ptr(y), _ptr(mask), ctypes.c_float(scale),
                                   ctypes.c_longlong(b * n * sq), sq, sk, n, b, mode, None) == 0
    z = x * scale
    if mode == 1:
        z = z.masked_fill(mask.bool(), -10000.0)
    if mode == 2:
        qi = torch.arange(sq)[:, None] + (sk - sq)
        z = z.masked_fill(torch.arange(sk)[None, :] > qi, float("-inf"))
    ref = torch.softmax(z, -1)
    if mode == 1:
        ref[0, :, 1] = 0.0
    assert (y - ref).abs().max().item() < 1e-6
    dy = torch.randn_like(x)
    ref_dx = scale * (ref * dy - ref * (ref * dy).sum(-1, keepdim=True))
    g = dy.clone()
    yy = ref.contiguous()
    assert kernels.mlb_softmax_bwd(DT[x.dtype], _ptr(g), _ptr(yy), ctypes.c_float(scale),
                                   ctypes.c_longlong(b * n * sq), sk, None) == 0
    assert (g - ref_dx).abs().max().item() < 1e-5


@pytest.mark.parametrize("rms", [True, False])
@pytest.mark.parametrize("dtype,H", [(torch.float32, 264), (torch.bfloat16, 1024)])
def test_norm_kernels_on_cpu_threads(kernels, rms, dtype, H):
    """csrc/norm.cu: RMSNorm / LayerNorm forward with the fused residual add, and the backward (dx with the residual
    gradient, per-part dw / db partial sums + the column-sum kernel)."""
    torch.manual_seed(2)
    rows, eps = 7, 1e-5
    x, res, w, bias = torch.randn(rows, H), torch.randn(rows, H), torch.rand(H) + 0.5, torch.randn(H)
    xd, rd, wd, bd = (t.to(dtype) for t in (x, res, w, bias))
    y, res_out = torch.empty_like(xd), torch.empty_like(xd)
    mean, rstd = torch.zeros(rows), torch.zeros(rows)
    assert kernels.mlb_norm_fwd(DT[dtype], _ptr(xd), _ptr(rd), _ptr(wd), _ptr(None if rms else bd), _ptr(y),
                                _ptr(res_out), _ptr(None if rms else mean), _ptr(rstd), rows, H, ctypes.c_float(eps),
                                int(rms), None) == 0
    xin = (xd.float() + rd.float()).to(dtype)                    # the residual stream is stored in `dtype`
    assert torch.equal(res_out, xin)
    xf = xin.float().requires_grad_()
    wf, bf = wd.float().requires_grad_(), bd.float().requires_grad_()
    if rms:
        ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * wf
    else:
        ref = torch.nn.functional.layer_norm(xf, (H,), wf, bf, eps)
    tol = 3e-2 if dtype == torch.bfloat16 else 2e-5
    assert (y.float() - ref).abs().max().item() < tol
    dy, dres = torch.randn(rows, H).to(dtype), torch.randn(rows, H).to(dtype)
    ref.backward(dy.float())
    parts = min(rows, 296)
    ws = torch.zeros(2 * 296 * H)
    dx, dw, db = torch.empty_like(xd), torch.empty_like(wd), torch.empty_like(wd)
    assert kernels.mlb_norm_bwd(DT[dtype], _ptr(dy), _ptr(xin), _ptr(wd), _ptr(None if rms else mean), _ptr(rstd),
                                _ptr(dres), _ptr(dx), _ptr(dw), _ptr(None if rms else db), _ptr(ws), parts, rows, H,
                                int(rms), None) == 0
    assert (dx.float() - (xf.grad + dres.float())).abs().max().item() < (6e-2 if dtype == torch.bfloat16 else 1e-4)
    assert (dw.float() - wf.grad).abs().max().item() < (1e-1 if dtype == torch.bfloat16 else 1e-4)
    if not rms:
        assert (db.float() - bf.grad).abs().max().item() < (1e-1 if dtype == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("kind,act", [(3, torch.nn.functional.silu), (1, torch.nn.functional.gelu), (2, torch.relu),
                                      (0, lambda z: z)])
def test_glu_kernels_on_cpu_threads(kernels, kind, act):
    """csrc/elementwise.cu: the GLU family forward / backward on the [rows, 2 F] projection output."""
    torch.manual_seed(3)
    rows, F_ = 3, 72
    x = torch.randn(rows, 2 * F_, requires_grad=True)
    a, g = x.chunk(2, dim=-1)
    ref = a * act(g)
    ref_alt = act(a) * g                      # (whichever half the kernel treats as the gate)
    y = torch.empty(rows, F_)
    assert kernels.mlb_glu_fwd(DT[torch.float32], _ptr(x.detach()), _ptr(y), ctypes.c_longlong(rows), F_, kind, None) == 0
    use = ref if (y - ref).abs().max() < (y - ref_alt).abs().max() else ref_alt
    assert (y - use).abs().max().item() < 1e-5
    dy = torch.randn(rows, F_)
    use.backward(dy)
    dx = torch.empty(rows, 2 * F_)
    assert kernels.mlb_glu_bwd(DT[torch.float32], _ptr(dy), _ptr(x.detach()), _ptr(dx), ctypes.c_longlong(rows), F_, kind,
                               None) == 0
    assert (dx - x.grad).abs().max().item() < 1e-5
