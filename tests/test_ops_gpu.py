"""Numerics of every hand-written sm_100a kernel against a plain PyTorch fp32 reference of the same op."""
import math
import os

import pytest
import torch

from megatron_llm_b200 import ops
from megatron_llm_b200.ops import _ext

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, atol, rtol=0.0):
    err = (a.float() - b.float()).abs().max().item()
    scale = b.float().abs().max().item()
    assert err <= atol + rtol * scale, f"max err {err} (ref scale {scale})"


def test_extension_is_loaded():
    mod = _ext.load()
    assert hasattr(mod, "gemm") and hasattr(mod, "norm_fwd")


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 320), (1000, 264, 136), (4096, 1536, 4096)])
def test_gemm_nt(M, N, K):
    torch.manual_seed(0)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    n0 = ops.launches()
    out = ops.gemm_nt(a, b)
    assert ops.launches() == n0 + 1
    _close(out, a.float() @ b.float().t(), atol=0.0, rtol=6e-3)


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (520, 264, 200), (2048, 4096, 1376)])
def test_gemm_nn_and_tn(M, N, K):
    torch.manual_seed(1)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(K, N, device=DEV, dtype=torch.bfloat16)
    _close(ops.gemm_nn(a, b), a.float() @ b.float(), atol=0.0, rtol=6e-3)
    at = torch.randn(K, M, device=DEV, dtype=torch.bfloat16)
    _close(ops.gemm_tn(at, b), at.float().t() @ b.float(), atol=0.0, rtol=6e-3)
    base = torch.randn(M, N, device=DEV, dtype=torch.float32)
    acc = base.clone()
    ops.gemm_tn(at, b, out=acc, accumulate=True)
    _close(acc, base + at.float().t() @ b.float(), atol=1e-3, rtol=1e-5)


@pytest.mark.parametrize("M,N,K", [(256, 512, 320), (4096, 1536, 4096)])      # 1-CTA and 2-CTA (cta_group::2) kernels
def test_gemm_fp16_operands(M, N, K):
    """The same tcgen05 kernels with IEEE fp16 operands (``kind::f16`` a/b format bits, fp16 epilogue packing): NT, NN,
    TN with fp16 / fp32 / fp32-accumulate outputs.  fp16 has 3 more mantissa bits than bf16: tighter bound."""
    torch.manual_seed(4)
    h = torch.float16
    a = torch.randn(M, K, device=DEV, dtype=h)
    b = torch.randn(N, K, device=DEV, dtype=h)
    n0 = ops.launches()
    out = ops.gemm_nt(a, b)
    assert ops.launches() == n0 + 1 and out.dtype == h
    _close(out, a.float() @ b.float().t(), atol=0.0, rtol=1.5e-3)
    bn = torch.randn(K, N, device=DEV, dtype=h)
    _close(ops.gemm_nn(a, bn), a.float() @ bn.float(), atol=0.0, rtol=1.5e-3)
    at = torch.randn(K, M, device=DEV, dtype=h)
    _close(ops.gemm_tn(at, bn), at.float().t() @ bn.float(), atol=0.0, rtol=1.5e-3)
    base = torch.randn(M, N, device=DEV, dtype=torch.float32)
    acc = base.clone()
    ops.gemm_tn(at, bn, out=acc, accumulate=True)
    _close(acc, base + at.float().t() @ bn.float(), atol=1e-3, rtol=1e-5)


def test_gemm_strided_views():
    torch.manual_seed(2)
    big = torch.randn(512, 768, device=DEV, dtype=torch.bfloat16)
    a = big[:, 128:128 + 256]           # row stride 768, 16B aligned
    b = torch.randn(320, 256, device=DEV, dtype=torch.bfloat16)
    _close(ops.gemm_nt(a, b), a.float() @ b.float().t(), atol=0.0, rtol=6e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("H", [256, 4096, 8192])
@pytest.mark.parametrize("rms", [True, False])
def test_norm_fwd_bwd(dtype, H, rms):
    torch.manual_seed(3)
    rows = 37
    x = torch.randn(rows, H, device=DEV, dtype=dtype)
    res = torch.randn(rows, H, device=DEV, dtype=dtype)
    w = (1 + 0.1 * torch.randn(H, device=DEV)).to(dtype)
    b = (0.1 * torch.randn(H, device=DEV)).to(dtype)
    for use_res in (False, True):
        xi, ri, wi, bi = (t.clone().requires_grad_(True) for t in (x, res, w, b))
        if rms:
            out = ops.rmsnorm(xi, wi, 1e-5, residual=ri if use_res else None)
        else:
            out = ops.layernorm(xi, wi, bi, 1e-5, residual=ri if use_res else None)
        y, hid = out if use_res else (out, None)
        # fp32 oracle
        xr, rr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, res, w, b))
        hin = (xr + rr).to(dtype).float() if use_res else xr
        if use_res:
            hin = ((xr + rr) - (xr + rr).detach() + (xr + rr).detach().to(dtype).float())
        if rms:
            yr = hin * torch.rsqrt(hin.pow(2).mean(-1, keepdim=True) + 1e-5) * wr
        else:
            yr = torch.nn.functional.layer_norm(hin, (H,), wr, br, 1e-5)
        tol = 2e-2 if dtype != torch.float32 else 1e-4
        _close(y, yr, atol=tol, rtol=tol)
        dy = torch.randn_like(y)
        dh = torch.randn_like(y)
        if use_res:
            torch.autograd.backward([y, hid], [dy, dh])
            torch.autograd.backward([yr, hin], [dy.float(), dh.float()])
        else:
            y.backward(dy)
            yr.backward(dy.float())
        _close(xi.grad, xr.grad, atol=tol * 4, rtol=tol)
        _close(wi.grad, wr.grad, atol=0.0, rtol=3e-2 if dtype != torch.float32 else 1e-4)
        if not rms:
            _close(bi.grad, br.grad, atol=0.0, rtol=3e-2 if dtype != torch.float32 else 1e-4)
        if use_res:
            _close(ri.grad, rr.grad, atol=tol * 4, rtol=tol)


@pytest.mark.parametrize("g", [1, 4])
def test_rope_matches_complex_reference(g):
    from megatron_llm_b200.models.positional_embeddings import apply_rotary_emb, precompute_freqs_cis
    torch.manual_seed(4)
    s, b, nkv, hn = 64, 2, 3, 128
    mixed = torch.randn(s, b, nkv * (g + 2) * hn, device=DEV, dtype=torch.bfloat16)
    ref = mixed.clone().view(s, b, nkv, g + 2, hn)
    table = ops.rope_table(hn, 128, theta=10000.0, scaling_factor=2.0, device=DEV)
    pos = torch.randint(0, 128, (b, s), device=DEV)
    src = mixed.clone().requires_grad_(True)
    out = ops.rope_qkv_(src.clone(), nkv, g + 2, hn, table, pos, 0).view(s, b, nkv, g + 2, hn)
    fc = precompute_freqs_cis(hn, 128, 10000.0, 2.0)
    q = ref[:, :, :, :g].reshape(s, b, nkv * g, hn)
    k = ref[:, :, :, g]
    qr, kr = apply_rotary_emb(q, k, fc, position_ids=pos)
    _close(out[:, :, :, :g].reshape(s, b, nkv * g, hn), qr, atol=2e-2)
    _close(out[:, :, :, g], kr, atol=2e-2)
    assert torch.equal(out[:, :, :, g + 1], ref[:, :, :, g + 1])   # V untouched
    # backward = inverse rotation: rotating the grad back must give orthogonal-transform consistency
    w = torch.randn_like(out)
    (ops.rope_qkv_(src.clone(), nkv, g + 2, hn, table, pos, 0).view(s, b, nkv, g + 2, hn) * w).sum().backward()
    assert torch.isfinite(src.grad).all()
    n1 = src.grad.view(s, b, nkv, g + 2, hn)[:, :, :, :g + 1].float().norm()
    n2 = w[:, :, :, :g + 1].float().norm()
    assert abs(n1 - n2) / n2 < 1e-2


@pytest.mark.parametrize("kind", ["swiglu", "geglu", "reglu", "liglu"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_glu(kind, dtype):
    torch.manual_seed(5)
    x = torch.randn(33, 2 * 1376, device=DEV, dtype=dtype)
    xi = x.clone().requires_grad_(True)
    y = ops.glu(xi, kind)
    xr = x.float().requires_grad_(True)
    x1, x2 = xr.chunk(2, -1)
    act = {"swiglu": torch.nn.functional.silu, "geglu": torch.nn.functional.gelu, "reglu": torch.relu,
           "liglu": lambda z: z}[kind]
    yr = x1 * act(x2)
    tol = 3e-2 if dtype == torch.bfloat16 else 1e-5
    _close(y, yr, atol=tol, rtol=tol)
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy.float())
    _close(xi.grad, xr.grad, atol=tol * 2, rtol=tol)


@pytest.mark.parametrize("approx", [False, True])
def test_gelu_with_bias(approx):
    torch.manual_seed(6)
    x = torch.randn(17, 512, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(512, device=DEV, dtype=torch.bfloat16)
    xi, bi = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ops.gelu(xi, bi, approximate=approx)
    xr, br = x.float().requires_grad_(True), b.float().requires_grad_(True)
    yr = torch.nn.functional.gelu(xr + br, approximate="tanh" if approx else "none")
    _close(y, yr, atol=3e-2)
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy.float())
    _close(xi.grad, xr.grad, atol=5e-2)
    _close(bi.grad, br.grad, atol=0.0, rtol=5e-2)


@pytest.mark.parametrize("V", [32000, 4000, 50257 + 7])
def test_cross_entropy_kernel(V):
    from megatron_llm_b200.parallel.cross_entropy import vocab_parallel_cross_entropy
    torch.manual_seed(7)
    T = 19
    Vp = (V + 7) // 8 * 8
    logits = (torch.randn(T, Vp, device=DEV) * 4).to(torch.bfloat16)
    tgt = torch.randint(0, Vp, (T,), device=DEV)
    li = logits.clone().requires_grad_(True)
    loss = vocab_parallel_cross_entropy(li, tgt)
    lr = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, tgt, reduction="none")
    _close(loss, ref, atol=2e-3, rtol=1e-4)
    w = torch.rand(T, device=DEV)
    (loss * w).sum().backward()
    (ref * w).sum().backward()
    _close(li.grad, lr.grad, atol=4e-3)


def test_softmax_family():
    torch.manual_seed(8)
    b, np_, sq, sk = 2, 3, 40, 40
    x = torch.randn(b, np_, sq, sk, device=DEV, dtype=torch.bfloat16)
    mask = torch.rand(b, 1, sq, sk, device=DEV) < 0.3
    for mode in ("plain", "mask", "causal"):
        xi = x.clone().requires_grad_(True)
        xr = x.float().requires_grad_(True)
        if mode == "plain":
            y = ops.scaled_softmax(xi, 0.5)
            yr = torch.softmax(xr * 0.5, -1)
        elif mode == "mask":
            y = ops.scaled_masked_softmax(xi, mask, 0.5)
            yr = torch.softmax((xr * 0.5).masked_fill(mask, -10000.0), -1)
        else:
            y = ops.scaled_upper_triang_masked_softmax(xi, 0.5)
            cm = torch.ones(sq, sk, device=DEV, dtype=torch.bool).tril()
            yr = torch.softmax((xr * 0.5).masked_fill(~cm, float("-inf")), -1)
        _close(y, yr, atol=1e-2)
        dy = torch.randn_like(y)
        y.backward(dy)
        yr.backward(dy.float())
        _close(xi.grad, xr.grad, atol=2e-2)


def test_flat_adamw_and_norm_match_torch():
    torch.manual_seed(9)
    mod = _ext.load()
    n = 100_003 // 8 * 8
    seg_start = torch.tensor([0, 4096, 40_000, n], dtype=torch.int64, device=DEV)
    seg_wd = torch.tensor([0.1, 0.0, 0.1], device=DEV)
    seg_w = torch.tensor([1.0, 0.0, 1.0], device=DEV)
    p = torch.randn(n, device=DEV)
    g = torch.randn(n, device=DEV) * 0.1
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    p16 = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    ws = torch.zeros(148 * 8, device=DEV)
    tot = torch.zeros(1, device=DEV)
    mod.sqnorm_flat(g, 0, seg_start, seg_w, ws, tot, False)
    ref_sq = g[:4096].pow(2).sum() + g[40_000:].pow(2).sum()
    assert abs(tot.item() - ref_sq.item()) / ref_sq.item() < 1e-5
    nrm, coef, inf = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV), torch.zeros(1, device=DEV, dtype=torch.int32)
    mod.clip_coef(tot, 1.0, nrm, coef, inf, 1.0)
    assert abs(nrm.item() - ref_sq.sqrt().item()) < 1e-3 and inf.item() == 0
    pr, mr, vr = p.clone(), m.clone(), v.clone()
    lr, b1, b2, eps = 1e-2, 0.9, 0.95, 1e-8
    for step in (1, 2, 3):
        mod.adamw_flat(p, g, m, v, p16, 0, seg_start, seg_wd, None, lr, b1, b2, eps, 1 - b1 ** step, 1 - b2 ** step,
                       coef, inf, [])
        gg = g * coef
        mr = b1 * mr + (1 - b1) * gg
        vr = b2 * vr + (1 - b2) * gg * gg
        wd = torch.repeat_interleave(seg_wd, seg_start[1:] - seg_start[:-1])
        pr = pr - lr * ((mr / (1 - b1 ** step)) / ((vr / (1 - b2 ** step)).sqrt() + eps) + wd * pr)
    _close(p, pr, atol=1e-5)
    _close(p16, pr, atol=2e-2)
    # skip flag turns the step into a no-op
    inf.fill_(1)
    before = p.clone()
    mod.adamw_flat(p, g, m, v, p16, 0, seg_start, seg_wd, None, lr, b1, b2, eps, 0.5, 0.5, coef, inf, [])
    assert torch.equal(p, before)
    # ZeRO-1 fused cast + all-gather: the 16-bit result goes to every listed buffer instead of `p16`
    inf.fill_(0)
    peers = [torch.zeros_like(p16) for _ in range(3)]
    p16.zero_()
    mod.adamw_flat(p, g, m, v, p16, 0, seg_start, seg_wd, None, lr, b1, b2, eps, 0.5, 0.5, coef, inf,
                   [t.data_ptr() for t in peers])
    for t in peers:
        assert torch.equal(t, p.to(p16.dtype))
    assert not p16.any()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_bias_dropout_add(dtype):
    torch.manual_seed(9)
    rows, F_ = 300, 1024
    x = torch.randn(rows, F_, device=DEV, dtype=dtype, requires_grad=True)
    b = torch.randn(F_, device=DEV, dtype=dtype, requires_grad=True)
    r = torch.randn(rows, F_, device=DEV, dtype=dtype, requires_grad=True)
    # p = 0 (and eval mode): exact bias + residual add
    out = ops.bias_dropout_add(x, b, r, 0.3, training=False)
    _close(out, (x.float() + b.float() + r.float()), atol=2e-2 if dtype != torch.float32 else 1e-6)
    # p > 0: inverted dropout of (x + b); survivors scaled by 1/(1-p); the backward mask is the forward mask.  The mask
    # only depends on (generator state, element index): probe it with x = 1, no bias, zero residual from the same state
    p = 0.25
    st = torch.cuda.get_rng_state()
    n0 = ops.launches()
    out = ops.bias_dropout_add(x, b, r, p, training=True)
    assert ops.launches() == n0 + 1
    torch.cuda.set_rng_state(st)
    kept = ops.bias_dropout_add(torch.ones_like(x), None, torch.zeros_like(r), p, training=True) != 0
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.01, frac
    want = torch.where(kept, (x.float() + b.float()) / (1 - p), torch.zeros_like(out, dtype=torch.float32)) + r.float()
    _close(out, want, atol=3e-2 if dtype != torch.float32 else 1e-5)
    do = torch.randn_like(out)
    out.backward(do)
    _close(x.grad, torch.where(kept, do.float() / (1 - p), torch.zeros_like(do, dtype=torch.float32)),
           atol=2e-2 if dtype != torch.float32 else 1e-5)
    _close(r.grad, do, atol=0.0)
    _close(b.grad, x.grad.float().sum(0), atol=0.5 if dtype != torch.float32 else 1e-3)
    # the seed follows the CUDA generator: same state -> same mask, next call -> a different one
    st = torch.cuda.get_rng_state()
    o1 = ops.bias_dropout_add(x, b, r, p, training=True)
    o2 = ops.bias_dropout_add(x, b, r, p, training=True)
    torch.cuda.set_rng_state(st)
    o3 = ops.bias_dropout_add(x, b, r, p, training=True)
    assert torch.equal(o1, o3) and not torch.equal(o1, o2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("sbh", [False, True])
def test_embedding_lookup_kernels(dtype, sbh):
    """Vocab-parallel gather (+ range mask + [s,b,h] transpose) and the scatter-add backward, dense and main_grad forms."""
    torch.manual_seed(13)
    b, s, H, V = 3, 64, 256, 512
    vocab_start, rows = 128, 256                                  # this "rank" owns ids [128, 384)
    ids = torch.randint(0, V, (b, s), device=DEV)
    ids[0, :8] = 200                                              # duplicates: their gradients must add up
    w = torch.randn(rows, H, device=DEV, dtype=dtype, requires_grad=True)
    out = ops.embedding_lookup(ids, w, vocab_start, sbh=sbh)
    local = ids - vocab_start
    mask = (local < 0) | (local >= rows)
    ref = torch.nn.functional.embedding(local.masked_fill(mask, 0), w.detach().float())
    ref = ref.masked_fill(mask.unsqueeze(-1), 0.0)
    ref = ref.transpose(0, 1).contiguous() if sbh else ref
    assert out.shape == ref.shape and torch.equal(out.float(), ref.to(dtype).float())
    do = torch.randn_like(out)
    out.backward(do)
    wr = w.detach().float().requires_grad_(True)
    o2 = torch.nn.functional.embedding(local.masked_fill(mask, 0), wr).masked_fill(mask.unsqueeze(-1), 0.0)
    (o2.transpose(0, 1) if sbh else o2).backward(do.float())
    _close(w.grad, wr.grad, atol=3e-2 if dtype == torch.bfloat16 else 1e-4)
    # fused form: accumulate into an fp32 main_grad, no .grad tensor
    w2 = w.detach().clone().requires_grad_(True)
    w2.main_grad = torch.ones(rows, H, device=DEV, dtype=torch.float32)
    ops.embedding_lookup(ids, w2, vocab_start, sbh=sbh, accumulate_into_main_grad=True).backward(do)
    assert w2.grad is None
    _close(w2.main_grad, 1.0 + wr.grad, atol=1e-4)


def test_accumulate_kernel():
    x = torch.randn(12345, device=DEV, dtype=torch.bfloat16)
    y = torch.randn(12345, device=DEV)
    ref = y + x.float()
    ops.accumulate_(y, x)
    _close(y, ref, atol=1e-6)


@pytest.mark.parametrize("nq,nkv", [(8, 8), (8, 2)])
def test_flash_attention_matches_reference(nq, nkv):
    from megatron_llm_b200.ops.attention import attention_reference
    torch.manual_seed(10)
    b, s, hn = 2, 256, 128
    q = torch.randn(b, s, nq, hn, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(b, s, nkv, hn, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(b, s, nkv, hn, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    out = ops.flash_attention(q, k, v, causal=True)
    ref = attention_reference(q.detach().float(), k.detach().float(), v.detach().float(), causal=True)
    _close(out, ref, atol=3e-2)
    out.sum().backward()
    assert torch.isfinite(q.grad).all() and torch.isfinite(k.grad).all()


@pytest.mark.parametrize("b,s,nq,nkv,window", [(2, 256, 8, 8, None), (2, 256, 8, 2, None), (1, 1024, 4, 1, 256),
                                               (1, 4096, 8, 2, None), (1, 4096, 4, 4, 1024)])
def test_attention_sm100_forward_and_all_grads_match_fp32(b, s, nq, nkv, window):
    """head_dim 128 tcgen05 kernels (fwd, dK/dV, dQ): output, dq, dk AND dv against an fp32 PyTorch attention, MHA / GQA
    / MQA, with and without a sliding window, up to the benchmark sequence length."""
    from megatron_llm_b200.ops import attention_sm100
    from megatron_llm_b200.ops.attention import attention_reference
    torch.manual_seed(12)
    hn = 128
    q, k, v = (torch.randn(b, s, n, hn, device=DEV, dtype=torch.bfloat16, requires_grad=True) for n in (nq, nkv, nkv))
    assert attention_sm100.supported(q, k, v, True, window, 0.0)
    out = attention_sm100.attention(q, k, v, True, window, None)
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ref = attention_reference(qr, kr, vr, causal=True, window=window)
    _close(out, ref, atol=3e-2)
    do = torch.randn_like(out)
    out.backward(do)
    ref.backward(do.float())
    for name, got, want in (("dq", q.grad, qr.grad), ("dk", k.grad, kr.grad), ("dv", v.grad, vr.grad)):
        err = (got.float() - want).abs().max().item()
        assert err <= 2e-2 * max(1.0, want.abs().max().item()), (name, err, want.abs().max().item())


@pytest.mark.parametrize("nq,nkv,window", [(4, 4, None), (8, 2, None), (8, 1, None), (4, 4, 128), (16, 1, None)])
def test_flash_attention_head_dim_64_matches_reference(nq, nkv, window):
    """Falcon / GPT-2 style heads: forward and all three input gradients against the fp32 reference."""
    from megatron_llm_b200.ops import attention_sm100
    from megatron_llm_b200.ops.attention import attention_reference
    torch.manual_seed(11)
    b, s, hn = 2, 512, 64
    q, k, v = (torch.randn(b, s, n, hn, device=DEV, dtype=torch.bfloat16, requires_grad=True) for n in (nq, nkv, nkv))
    assert attention_sm100.supported(q, k, v, True, window, 0.0)
    out = attention_sm100.attention(q, k, v, True, window, None)
    qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ref = attention_reference(qr, kr, vr, causal=True, window=window)
    _close(out, ref, atol=3e-2)
    do = torch.randn_like(out)
    out.backward(do)
    ref.backward(do.float())
    for got, want in ((q.grad, qr.grad), (k.grad, kr.grad), (v.grad, vr.grad)):
        assert (got.float() - want).abs().max() <= 3e-2 * max(1.0, want.abs().max().item())
