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
    so = host_build.build(["attention_decode.cu", "ce.cu", "softmax.cu", "norm.cu", "elementwise.cu", "optim.cu",
                           "embedding.cu"],
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


def test_flat_optimizer_kernels_on_cpu_threads(kernels):
    """csrc/optim.cu over a flat buffer of three parameters (segment table with per-parameter weight decay / lr
    multiplier, a ragged tail, a shard offset): AdamW with the device-side gradient scale and the bf16 write-back, the
    skip flag, the weighted squared norm and the clip coefficient."""
    torch.manual_seed(4)
    off = 5000                                             # this shard starts at element 5000 of the global buffer
    sizes = [3000, 1111, 2050]
    n = sum(sizes)
    seg_start = torch.tensor([off, off + 3000, off + 4111, off + n])
    wd, lrm, wgt = torch.tensor([0.1, 0.0, 0.05]), torch.tensor([1.0, 2.0, 0.5]), torch.tensor([1.0, 0.5, 0.0])
    p, g, m, v = torch.randn(n), torch.randn(n) * 4, torch.randn(n) * 0.1, torch.rand(n) * 0.1
    lr, b1, b2, eps, step, gscale = 1e-2, 0.9, 0.95, 1e-8, 3, torch.tensor([0.25])
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    seg_of = torch.repeat_interleave(torch.arange(3), torch.tensor(sizes))
    gg = g * gscale
    m_ref = b1 * m + (1 - b1) * gg
    v_ref = b2 * v + (1 - b2) * gg * gg
    p_ref = p - lr * lrm[seg_of] * ((m_ref / bc1) / ((v_ref / bc2).sqrt() + eps) + wd[seg_of] * p)
    # skip flag set: nothing moves
    pk, mk, vk, p16 = p.clone(), m.clone(), v.clone(), torch.zeros(n, dtype=torch.bfloat16)
    flag = torch.tensor([1], dtype=torch.int32)
    args = lambda fl: (_ptr(pk), _ptr(g), _ptr(mk), _ptr(vk), _ptr(p16), DT[torch.bfloat16], ctypes.c_longlong(n),
                       ctypes.c_longlong(off), _ptr(seg_start), _ptr(wd), _ptr(lrm), 3, ctypes.c_float(lr),
                       ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(eps), ctypes.c_float(bc1),
                       ctypes.c_float(bc2), _ptr(gscale), _ptr(fl), None, 0, None)
    assert kernels.mlb_adamw_flat(*args(flag)) == 0
    assert torch.equal(pk, p) and torch.equal(mk, m) and torch.equal(vk, v)
    flag.zero_()
    assert kernels.mlb_adamw_flat(*args(flag)) == 0
    assert torch.allclose(mk, m_ref, atol=1e-6) and torch.allclose(vk, v_ref, atol=1e-6)
    assert torch.allclose(pk, p_ref, atol=1e-6)
    assert torch.equal(p16, pk.bfloat16())
    # weighted squared norm (weights drop TP-duplicated parameters), accumulated onto a previous value
    ws, out = torch.zeros(148 * 8), torch.tensor([2.0])
    gb = g.bfloat16()
    assert kernels.mlb_sqnorm_flat(DT[torch.bfloat16], _ptr(gb), ctypes.c_longlong(n), ctypes.c_longlong(off),
                                   _ptr(seg_start), _ptr(wgt), 3, _ptr(ws), _ptr(out), 1, None) == 0
    ref = 2.0 + (wgt[seg_of] * gb.float() ** 2).sum()
    assert abs(out.item() - ref.item()) < 1e-3 * ref.item()
    # clip coefficient: min(1, max_norm / (norm + 1e-6)) * norm_scale, and the inf flag
    nrm, coef, inf = torch.zeros(1), torch.zeros(1), torch.tensor([7], dtype=torch.int32)
    tot = torch.tensor([400.0])
    assert kernels.mlb_clip_coef(_ptr(tot), ctypes.c_float(1.0), _ptr(nrm), _ptr(coef), _ptr(inf), ctypes.c_float(0.5),
                                 None) == 0
    assert abs(nrm.item() - 10.0) < 1e-5 and abs(coef.item() - 0.5 * (1.0 / (10.0 + 1e-6))) < 1e-7 and inf.item() == 0
    tot[0] = float("inf")
    assert kernels.mlb_clip_coef(_ptr(tot), ctypes.c_float(1.0), _ptr(nrm), _ptr(coef), _ptr(inf), ctypes.c_float(1.0),
                                 None) == 0
    assert inf.item() == 1


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bias_dropout_add_and_gelu_kernels_on_cpu_threads(kernels, dtype):
    """csrc/elementwise.cu: residual + dropout(x + bias) forward, its backward (the same mask regenerated from the
    seed), and bias + GeLU forward / backward."""
    torch.manual_seed(5)
    rows, F_, pdrop, seed = 6, 64, 0.3, 123456789
    x, bias, res = (torch.randn(*sh).to(dtype) for sh in ((rows, F_), (F_,), (rows, F_)))
    out = torch.empty_like(x)
    assert kernels.mlb_bias_dropout_add(DT[dtype], _ptr(x), _ptr(bias), _ptr(res), _ptr(out), ctypes.c_longlong(rows), F_,
                                        ctypes.c_float(pdrop), ctypes.c_ulonglong(seed), 0, None) == 0
    full = (x.float() + bias.float()) / (1 - pdrop) + res.float()
    kept = (out.float() - full).abs() < (5e-2 if dtype == torch.bfloat16 else 1e-5)
    dropped = (out.float() - res.float()).abs() < 1e-6
    assert (kept | dropped).all() and 0.15 < 1 - kept.float().mean().item() < 0.45
    dy, dx = torch.randn(rows, F_).to(dtype), torch.empty_like(x)
    assert kernels.mlb_bias_dropout_add(DT[dtype], _ptr(dy), None, None, _ptr(dx), ctypes.c_longlong(rows), F_,
                                        ctypes.c_float(pdrop), ctypes.c_ulonglong(seed), 1, None) == 0
    unambiguous = kept != dropped
    ref_dx = torch.where(kept, dy.float() / (1 - pdrop), torch.zeros(()))
    assert (dx.float() - ref_dx)[unambiguous].abs().max().item() < (5e-2 if dtype == torch.bfloat16 else 1e-5)
    # p = 0: plain bias + residual add
    assert kernels.mlb_bias_dropout_add(DT[dtype], _ptr(x), _ptr(bias), _ptr(res), _ptr(out), ctypes.c_longlong(rows), F_,
                                        ctypes.c_float(0.0), ctypes.c_ulonglong(seed), 0, None) == 0
    assert (out.float() - (x.float() + bias.float() + res.float())).abs().max().item() < (5e-2 if dtype == torch.bfloat16 else 1e-6)
    # bias + GeLU (exact and tanh approximation), forward and backward
    for approx in (0, 1):
        xg = (x.float() + bias.float()).requires_grad_()
        ref = torch.nn.functional.gelu(xg, approximate="tanh" if approx else "none")
        ref.backward(dy.float())
        y, dxg = torch.empty_like(x), torch.empty_like(x)
        assert kernels.mlb_gelu(DT[dtype], _ptr(x), _ptr(bias), None, _ptr(y), ctypes.c_longlong(rows), F_, approx, 0, None) == 0
        assert kernels.mlb_gelu(DT[dtype], _ptr(x), _ptr(bias), _ptr(dy), _ptr(dxg), ctypes.c_longlong(rows), F_, approx, 1, None) == 0
        tol = 5e-2 if dtype == torch.bfloat16 else 2e-3
        assert (y.float() - ref).abs().max().item() < tol and (dxg.float() - xg.grad).abs().max().item() < tol


@pytest.mark.parametrize("with_ids", [False, True])
def test_rope_kernel_on_cpu_threads(kernels, with_ids):
    """csrc/elementwise.cu rope_qkv: in-place rotation of the q heads and the k head of the packed QKV buffer (the v
    head untouched), from a position offset or explicit position ids, and the inverse rotation used in backward."""
    from megatron_llm_b200 import ops
    torch.manual_seed(6)
    s, b, nkv, hpg, hn = 5, 2, 2, 4, 16                    # hpg = 2 query heads + k + v per group
    table = ops.rope_table(hn, 64)
    mixed = torch.randn(s, b, nkv * hpg * hn)
    pos = torch.randint(0, 64, (b, s)) if with_ids else None
    ref = ops._rope_qkv_apply_(mixed.clone(), nkv, hpg, hn, table, pos, 0 if with_ids else 7, False)   # torch path
    got = mixed.clone()
    assert kernels.mlb_rope_qkv(DT[torch.float32], _ptr(got), _ptr(table), _ptr(pos), s * b, b, nkv, hpg, hn,
                                0 if with_ids else 7, 0, ctypes.c_longlong(nkv * hpg * hn), None) == 0
    assert (got - ref).abs().max().item() < 1e-5
    v_slice = got.view(s, b, nkv, hpg, hn)[:, :, :, -1]
    assert torch.equal(v_slice, mixed.view(s, b, nkv, hpg, hn)[:, :, :, -1])
    assert kernels.mlb_rope_qkv(DT[torch.float32], _ptr(got), _ptr(table), _ptr(pos), s * b, b, nkv, hpg, hn,
                                0 if with_ids else 7, 1, ctypes.c_longlong(nkv * hpg * hn), None) == 0
    assert (got - mixed).abs().max().item() < 1e-5          # inverse rotation restores the input


@pytest.mark.parametrize("sbh", [0, 1])
def test_embedding_kernels_on_cpu_threads(kernels, sbh):
    """csrc/embedding.cu: vocab-shard gather (rows of other shards give zeros, optional [s, b, h] output order) and the
    scatter-add backward into the fp32 main gradient (repeated ids accumulate)."""
    torch.manual_seed(7)
    b, s, H, start, rows_local = 3, 11, 40, 100, 50
    ids = torch.randint(start - 10, start + rows_local + 10, (b, s))
    ids[0, :4] = start + 7                                   # repeated id -> accumulation in backward
    w = torch.randn(rows_local, H).bfloat16()
    out = torch.full((s, b, H) if sbh else (b, s, H), float("nan")).bfloat16()
    assert kernels.mlb_embedding_fwd(DT[torch.bfloat16], _ptr(ids), _ptr(w), _ptr(out), b, s, H, ctypes.c_longlong(start),
                                     ctypes.c_longlong(rows_local), sbh, None) == 0
    local = ids - start
    owned = (local >= 0) & (local < rows_local)
    ref = torch.where(owned[..., None], w[local.clamp(0, rows_local - 1)].float(), torch.zeros(()))
    ref = ref.transpose(0, 1) if sbh else ref
    assert torch.equal(out.float(), ref)
    dout = torch.randn_like(out.float()).bfloat16()
    dw = torch.ones(rows_local, H)                           # accumulates into what is already there
    assert kernels.mlb_embedding_bwd(DT[torch.bfloat16], _ptr(ids), _ptr(dout), _ptr(dw), b, s, H, ctypes.c_longlong(start),
                                     ctypes.c_longlong(rows_local), sbh, None) == 0
    d = (dout.float().transpose(0, 1) if sbh else dout.float()).reshape(b * s, H)
    ref_dw = torch.ones(rows_local, H).index_add_(0, local.reshape(-1)[owned.reshape(-1)], d[owned.reshape(-1)])
    assert (dw - ref_dw).abs().max().item() < 1e-4


# ------------------------------------------------------------------------------------------------------------------
# shared-memory race check: the same host build under ThreadSanitizer
# ------------------------------------------------------------------------------------------------------------------

def test_simt_kernels_have_no_shared_memory_races(tmp_path):
    """One OS thread per CUDA thread makes a missing __syncthreads / __syncwarp a data race that ThreadSanitizer sees:
    decode attention, norms, cross-entropy, softmax, the flat norm reduction and the embedding scatter-add run clean,
    and a decode kernel with one __syncwarp removed on purpose is reported (so a clean run means something)."""
    import subprocess
    files = ["attention_decode.cu", "norm.cu", "ce.cu", "softmax.cu", "optim.cu", "embedding.cu"]
    exe = host_build.build_race_driver(files, str(tmp_path / "clean"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, r.stderr[-3000:]

    def drop_a_barrier(name, src):
        old = "      __syncwarp();\n      // ---- acc += P V"
        assert src.count(old) == 1
        return src.replace(old, "      // ---- acc += P V")
    exe = host_build.build_race_driver(["attention_decode.cu"], str(tmp_path / "mutant"), mutate=drop_a_barrier,
                                       defines=["RACE_DECODE_ONLY"])
    r = subprocess.run([exe, "decode"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "ThreadSanitizer: data race" in r.stderr


# ------------------------------------------------------------------------------------------------------------------
# peer-memory collectives: ranks = processes, symmetric memory = files every rank maps (tests/emu/comm_rank.py)
# ------------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def comm_lib(tmp_path_factory):
    return host_build.build(["comm.cu"], str(tmp_path_factory.mktemp("emu_comm")))


def _ranks(comm_lib, d, ranks, world, n, mode, epoch=1, ctas=2, timeout=300):
    import subprocess
    present = ",".join(str(r) for r in ranks)
    procs = [subprocess.Popen([sys.executable, os.path.join(EMU, "comm_rank.py"), comm_lib, str(d), str(r), str(world),
                               str(n), mode, str(epoch), str(ctas), present]) for r in ranks]
    return [p.wait(timeout=timeout) for p in procs]


def _symmetric_memory(d, world, n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    data = [rng.standard_normal(n).astype(np.float32) for _ in range(world)]
    for r in range(world):
        data[r].tofile(os.path.join(d, f"buf{r}.bin"))
        np.zeros(64, dtype=np.int32).tofile(os.path.join(d, f"pad{r}.bin"))
    return data


@pytest.mark.parametrize("mode,world", [("peer_ar", 2), ("peer_rs", 3), ("nvls_ar", 3), ("nvls_rs", 2)])
def test_dp_reduction_kernels_between_emulated_ranks(comm_lib, tmp_path, mode, world):
    """csrc/comm.cu: the two-shot DP gradient reduction (peer loads, and the NVLS form with the in-switch reduction
    emulated) between concurrent ranks: handshake in, reduce + 1/DP scale, all-gather / scatter, handshake out; then a
    second epoch on the same pads."""
    import numpy as np
    n = world * 4 * 700
    data = _symmetric_memory(tmp_path, world, n, seed=world)
    assert _ranks(comm_lib, tmp_path, range(world), world, n, mode, epoch=1) == [0] * world
    mean = sum(data) / world
    slice_ = n // world
    for r in range(world):
        got = np.fromfile(os.path.join(tmp_path, f"buf{r}.bin"), dtype=np.float32)
        pad = np.fromfile(os.path.join(tmp_path, f"pad{r}.bin"), dtype=np.int32)
        assert pad[32] == 0 and pad[40] == 0                         # no timeout, CTA counter reset
        if mode.endswith("_ar"):
            assert np.allclose(got, mean, atol=1e-6)
        else:                                                        # reduce-scatter: rank r owns slice r
            assert np.allclose(got[r * slice_:(r + 1) * slice_], mean[r * slice_:(r + 1) * slice_], atol=1e-6)
            other = (r + 1) % world
            assert np.array_equal(got[other * slice_:(other + 1) * slice_], data[r][other * slice_:(other + 1) * slice_])
    if mode.endswith("_ar"):                                         # next step: epoch 2 over the reduced buffers
        assert _ranks(comm_lib, tmp_path, range(world), world, n, mode, epoch=2) == [0] * world
        got = np.fromfile(os.path.join(tmp_path, "buf0.bin"), dtype=np.float32)
        assert np.allclose(got, mean, atol=1e-6)                     # the mean of equal copies is the copy


def test_lost_peer_times_out_instead_of_hanging(comm_lib, tmp_path):
    """A rank whose peer never arrives leaves its bounded spin, records the timeout in its pad (what
    ``symm.check_timeouts`` polls once per training step) and the kernel retires."""
    import numpy as np
    n = 2 * 4 * 64
    _symmetric_memory(tmp_path, 2, n, seed=9)
    assert _ranks(comm_lib, tmp_path, [0], 2, n, "barrier", ctas=1) == [0]
    pad = np.fromfile(os.path.join(tmp_path, "pad0.bin"), dtype=np.int32)
    assert pad[32] == 1
    # both present: the barrier completes without the flag
    _symmetric_memory(tmp_path, 2, n, seed=9)
    assert _ranks(comm_lib, tmp_path, [0, 1], 2, n, "barrier", ctas=1) == [0, 0]
    assert all(np.fromfile(os.path.join(tmp_path, f"pad{r}.bin"), dtype=np.int32)[32] == 0 for r in range(2))
