"""Hot-path operators.

Every op has two implementations with identical semantics:
  * a hand-written sm_100a kernel from ``csrc/`` (used for every CUDA tensor; a missing extension on a GPU
    box is a hard error, never a silent library fallback), and
  * a plain PyTorch reference (CPU/Gloo plumbing runs and the numerical oracle for the tests).

Autograd wrappers live here so models call ``ops.rmsnorm(x, w, eps)`` etc. and get the fused forward AND
backward kernels.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from . import _ext
from ._ext import cuda_ops_available, count as _count


def _C():
    return _ext.load()


def launches() -> int:
    return _ext.LAUNCHES


# =============================================================================================
# GEMM (tcgen05).  Layout convention follows torch.nn.Linear: weight is [out, in].
# =============================================================================================

_EPI_BF16, _EPI_F32_ACCUM, _EPI_F32, _EPI_BF16_ACCUM = 0, 1, 2, 3


def _tile_cfg(M: int, N: int) -> int:
    """block_n selector passed to the kernel launcher: 0 = 1-CTA kernel with the wave-quantisation heuristic,
    512 = 2-CTA (cta_group::2) 256x256 tiles.  MLB200_GEMM_2CTA=0/1 forces one or the other."""
    import os
    mode = os.environ.get("MLB200_GEMM_2CTA", "auto")
    if mode == "0":
        return 0
    if mode == "1":
        return 512
    # auto: the 2-CTA kernel halves the B-operand shared-memory / L2 traffic per SM and wins whenever there are enough
    # 256x256 tiles to keep the 74 SM pairs busy (measured: profiles/gemm_shapes_v2.jsonl)
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    return 512 if (M >= 256 and N >= 256 and tiles >= 48) else 0


def _gemm_ok(*ts) -> bool:
    """bf16 x bf16 or fp16 x fp16 CUDA operands (the tcgen05 ``kind::f16`` MMA takes either; fp32 models use the
    library)."""
    for t in ts:
        if not (t.is_cuda and t.dtype in (torch.bfloat16, torch.float16) and t.dtype == ts[0].dtype):
            return False
    return cuda_ops_available(ts[0])


def _rowmajor2d(t: torch.Tensor) -> torch.Tensor:
    """2-D view with unit inner stride and 16-byte aligned rows (copies only when it has to)."""
    if t.dim() != 2:
        t = t.reshape(-1, t.size(-1))
    if t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def gemm_nt(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, comm=None,
            sms: int = 0) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T   (forward of a Linear: x @ W^T)."""
    if not _gemm_ok(a, b) or a.size(1) % 8 or b.size(0) % 8:
        r = a @ b.t()
        if out is not None:
            out.copy_(r)
            return out
        return r
    a, b = _rowmajor2d(a), _rowmajor2d(b)
    M, K = a.shape
    N = b.size(0)
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _C().gemm(a, b, out, M, N, K, a.stride(0), b.stride(0), out.stride(0), False, False, _EPI_BF16, _tile_cfg(M, N), comm, sms)
    _count()
    return out


def gemm_nn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, comm=None,
            sms: int = 0) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[K,N]   (dgrad: dY @ W)."""
    if not _gemm_ok(a, b) or a.size(1) % 8 or b.size(1) % 8:
        r = a @ b
        if out is not None:
            out.copy_(r)
            return out
        return r
    a, b = _rowmajor2d(a), _rowmajor2d(b)
    M, K = a.shape
    N = b.size(1)
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
    _C().gemm(a, b, out, M, N, K, a.stride(0), b.stride(0), out.stride(0), False, True, _EPI_BF16, _tile_cfg(M, N), comm, sms)
    _count()
    return out


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False,
            sms: int = 0) -> torch.Tensor:
    """out[M,N] (+)= a[K,M]^T @ b[K,N]   (wgrad: dY^T @ X; ``out`` may be the fp32 main_grad)."""
    M, N, K = a.size(1), b.size(1), a.size(0)
    if not _gemm_ok(a, b) or M % 8 or N % 8 or (out is not None and out.dtype not in (torch.float32, a.dtype)):
        r = a.t().float() @ b.float() if (out is not None and out.dtype == torch.float32) else a.t() @ b
        if out is None:
            return r
        if accumulate:
            out.add_(r.to(out.dtype))
        else:
            out.copy_(r)
        return out
    a, b = _rowmajor2d(a), _rowmajor2d(b)
    if out is None:
        out = torch.empty((M, N), dtype=a.dtype, device=a.device)
        accumulate = False
    if out.dtype == torch.float32:
        epi = _EPI_F32_ACCUM if accumulate else _EPI_F32
    else:
        epi = _EPI_BF16_ACCUM if accumulate else _EPI_BF16
    _C().gemm(a, b, out, M, N, K, a.stride(0), b.stride(0), out.stride(0), True, True, epi, _tile_cfg(M, N), None, sms)
    _count()
    return out


# =============================================================================================
# RMSNorm / LayerNorm (optionally fused with the residual add that precedes them)
# =============================================================================================

_NORM_PARTS = 148 * 2
_norm_ws = {}


def _norm_workspace(device, H):
    key = (device, H)
    ws = _norm_ws.get(key)
    if ws is None:
        ws = torch.empty(2 * _NORM_PARTS * H, dtype=torch.float32, device=device)
        _norm_ws[key] = ws
    return ws


def _norm_kernel_ok(x, w):
    H = x.size(-1)
    return (x.is_cuda and x.dtype == w.dtype and x.dtype in (torch.bfloat16, torch.float16, torch.float32)
            and H % 8 == 0 and H <= 8192 and cuda_ops_available(x))


class _NormFn(torch.autograd.Function):
    """y = norm(x [+ residual]); returns (y, x+residual) when a residual is given."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps, rms):
        xc = x.contiguous()
        H = xc.size(-1)
        ctx.rms, ctx.has_res, ctx.has_bias = rms, residual is not None, bias is not None
        if _norm_kernel_ok(xc, weight):
            y = torch.empty_like(xc)
            rows = xc.numel() // H
            rstd = torch.empty(rows, dtype=torch.float32, device=xc.device)
            mean = None if rms else torch.empty(rows, dtype=torch.float32, device=xc.device)
            res_out = torch.empty_like(xc) if residual is not None else None
            _C().norm_fwd(xc, residual.contiguous() if residual is not None else None, weight, bias, y, res_out,
                          mean, rstd, eps, rms)
            _count()
            xin = res_out if residual is not None else xc
            ctx.kernel = True
        else:
            xin = xc if residual is None else (xc + residual)
            xf = xin.float()
            # fp32 residual stream with half-precision parameters (--fp32_residual_connection): the normed
            # activation feeds the GEMMs, so it takes the parameter dtype (apex's "mixed" norm does the same)
            ydt = weight.dtype if (xc.dtype == torch.float32 and weight.dtype != torch.float32) else xc.dtype
            if rms:
                mean = None
                rstd = torch.rsqrt(xf.pow(2).mean(-1) + eps).reshape(-1)
                y = (xf * rstd.view(*xf.shape[:-1], 1) * weight.float()).to(ydt)
            else:
                mean = xf.mean(-1).reshape(-1)
                var = xf.var(-1, unbiased=False).reshape(-1)
                rstd = torch.rsqrt(var + eps)
                y = ((xf - mean.view(*xf.shape[:-1], 1)) * rstd.view(*xf.shape[:-1], 1) * weight.float())
                if bias is not None:
                    y = y + bias.float()
                y = y.to(ydt)
            ctx.kernel = False
        ctx.save_for_backward(xin, weight, mean, rstd)
        if residual is not None:
            return y, xin
        return y

    @staticmethod
    def backward(ctx, dy, dres=None):
        xin, weight, mean, rstd = ctx.saved_tensors
        H = xin.size(-1)
        dy = dy.contiguous()
        if not ctx.has_res:
            dres = None
        if ctx.kernel:
            dx = torch.empty_like(xin)
            dw = torch.empty_like(weight)
            db = torch.empty_like(weight) if (ctx.has_bias and not ctx.rms) else None
            rows = xin.numel() // H
            parts = min(rows, _NORM_PARTS)
            _C().norm_bwd(dy, xin, weight, mean, rstd, dres.contiguous() if dres is not None else None, dx, dw, db,
                          _norm_workspace(xin.device, H), parts, ctx.rms)
            _count(2)
        else:
            xf, g = xin.float(), dy.float() * weight.float()
            r = rstd.view(*xf.shape[:-1], 1)
            xh = xf * r if ctx.rms else (xf - mean.view(*xf.shape[:-1], 1)) * r
            c2 = (g * xh).mean(-1, keepdim=True)
            c1 = 0 if ctx.rms else g.mean(-1, keepdim=True)
            dxf = r * (g - c1 - xh * c2)
            if dres is not None:
                dxf = dxf + dres.float()
            dx = dxf.to(xin.dtype)
            dw = (dy.float() * xh).reshape(-1, H).sum(0).to(weight.dtype)
            db = dy.float().reshape(-1, H).sum(0).to(weight.dtype) if (ctx.has_bias and not ctx.rms) else None
        if ctx.has_res:
            return dx, dx, dw, db, None, None
        return dx, None, dw, db, None, None


def rmsnorm(x, weight, eps=1e-5, residual=None):
    """RMSNorm; with ``residual`` returns (norm(x+residual), x+residual)."""
    return _NormFn.apply(x, residual, weight, None, eps, True)


def layernorm(x, weight, bias, eps=1e-5, residual=None):
    return _NormFn.apply(x, residual, weight, bias, eps, False)


# =============================================================================================
# RoPE (interleaved-pair convention), in place
# =============================================================================================

def rope_table(dim: int, end: int, theta: float = 10000.0, scaling_factor: float = 1.0,
               device=None) -> torch.Tensor:
    """[end, dim/2, 2] fp32 (cos, sin); positions are divided by ``scaling_factor`` (linear RoPE scaling).
    Same frequencies as the reference's complex ``freqs_cis`` (positional_embeddings.py:7-13)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    t = torch.arange(end, dtype=torch.float32) / scaling_factor
    ang = torch.outer(t, freqs)
    tab = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
    return tab.to(device) if device is not None else tab


def _rope_ref(x, table, pos, inverse):
    """x [..., hn] with pos broadcastable to x.shape[:-1]; returns rotated copy."""
    cs = table[pos]  # [..., hn/2, 2]
    c, s = cs[..., 0], cs[..., 1]
    if inverse:
        s = -s
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    x0, x1 = xf[..., 0], xf[..., 1]
    out = torch.stack([x0 * c - x1 * s, x0 * s + x1 * c], dim=-1)
    return out.reshape(x.shape).to(x.dtype)


def _rope_qkv_apply_(mixed, nkv, hpg, hn, table, position_ids, pos_offset, inverse):
    """mixed [s, b, nkv*hpg*hn] viewed as [s, b, nkv, hpg, hn]: rotate the q heads and the k head in place."""
    s, b = mixed.shape[:2]
    assert table.dtype == torch.float32, "RoPE (cos, sin) table must stay fp32"
    if cuda_ops_available(mixed) and mixed.is_contiguous() and hn % 8 == 0:
        if position_ids is not None and not position_ids.is_contiguous():
            position_ids = position_ids.contiguous()
        _C().rope_qkv(mixed, table, position_ids, s * b, b, nkv, hpg, hn, pos_offset, inverse, nkv * hpg * hn)
        _count()
        return mixed
    qkv = mixed.view(s, b, nkv, hpg, hn)
    if position_ids is not None:
        pos = position_ids.t().reshape(s, b, 1, 1)
    else:
        pos = (torch.arange(s, device=mixed.device) + pos_offset).view(s, 1, 1, 1)
    rot = qkv[:, :, :, :-1]
    rot.copy_(_rope_ref(rot, table, pos.expand(s, b, nkv, hpg - 1), inverse))
    return mixed


class _RopeQKVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mixed, nkv, hpg, hn, table, position_ids, pos_offset):
        ctx.save_for_backward(table, position_ids)
        ctx.meta = (nkv, hpg, hn, pos_offset)
        ctx.mark_dirty(mixed)
        return _rope_qkv_apply_(mixed, nkv, hpg, hn, table, position_ids, pos_offset, False)

    @staticmethod
    def backward(ctx, g):
        table, position_ids = ctx.saved_tensors
        nkv, hpg, hn, pos_offset = ctx.meta
        g = g.contiguous()
        return _rope_qkv_apply_(g, nkv, hpg, hn, table, position_ids, pos_offset, True), None, None, None, None, \
            None, None


def rope_qkv_(mixed, nkv, heads_per_group, hn, table, position_ids=None, pos_offset=0):
    """In-place RoPE on the packed QKV GEMM output ``mixed`` [s, b, nkv*(g+2)*hn] (differentiable).  ``mixed`` must
    be a base tensor (the Linear autograd function returns one), not a view."""
    return _RopeQKVFn.apply(mixed, nkv, heads_per_group, hn, table, position_ids, pos_offset)


# =============================================================================================
# GLU family / GeLU
# =============================================================================================

_GLU_KINDS = {"liglu": 0, "geglu": 1, "reglu": 2, "swiglu": 3}


def _glu_act(kind, z):
    if kind == 0:
        return z
    if kind == 1:
        return F.gelu(z)
    if kind == 2:
        return F.relu(z)
    return F.silu(z)


class _GluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        ctx.kind = kind
        xc = x.contiguous()
        ctx.save_for_backward(xc)
        Fh = xc.size(-1) // 2
        if cuda_ops_available(xc) and Fh % 8 == 0 and xc.dtype in (torch.bfloat16, torch.float16, torch.float32):
            y = torch.empty(xc.shape[:-1] + (Fh,), dtype=xc.dtype, device=xc.device)
            _C().glu_fwd(xc, y, kind)
            _count()
            return y
        x1, x2 = xc.chunk(2, dim=-1)
        return (x1.float() * _glu_act(kind, x2.float())).to(xc.dtype)

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        dy = dy.contiguous()
        Fh = xc.size(-1) // 2
        if cuda_ops_available(xc) and Fh % 8 == 0 and xc.dtype in (torch.bfloat16, torch.float16, torch.float32):
            dx = torch.empty_like(xc)
            _C().glu_bwd(dy, xc, dx, ctx.kind)
            _count()
            return dx, None
        with torch.enable_grad():
            xd = xc.detach().float().requires_grad_(True)
            x1, x2 = xd.chunk(2, dim=-1)
            y = x1 * _glu_act(ctx.kind, x2)
            (dx,) = torch.autograd.grad(y, xd, dy.float())
        return dx.to(xc.dtype), None


def glu(x, kind: str):
    """x[..., 2F] -> x1 * act(x2) with x1 = first half (up), x2 = second half (gate)."""
    return _GluFn.apply(x, _GLU_KINDS[kind])


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, approx):
        xc = x.contiguous()
        ctx.approx = approx
        ctx.save_for_backward(xc, bias)
        if cuda_ops_available(xc) and xc.size(-1) % 8 == 0 and xc.dtype in (torch.bfloat16, torch.float16, torch.float32):
            y = torch.empty_like(xc)
            _C().gelu(xc, bias, None, y, approx, False)
            _count()
            return y
        z = xc.float() + (bias.float() if bias is not None else 0)
        return F.gelu(z, approximate="tanh" if approx else "none").to(xc.dtype)

    @staticmethod
    def backward(ctx, dy):
        xc, bias = ctx.saved_tensors
        dy = dy.contiguous()
        if cuda_ops_available(xc) and xc.size(-1) % 8 == 0 and xc.dtype in (torch.bfloat16, torch.float16, torch.float32):
            dx = torch.empty_like(xc)
            _C().gelu(xc, bias, dy, dx, ctx.approx, True)
            _count()
        else:
            with torch.enable_grad():
                z = (xc.float() + (bias.float() if bias is not None else 0)).detach().requires_grad_(True)
                y = F.gelu(z, approximate="tanh" if ctx.approx else "none")
                (dx,) = torch.autograd.grad(y, z, dy.float())
            dx = dx.to(xc.dtype)
        db = dx.reshape(-1, dx.size(-1)).sum(0) if bias is not None else None
        return dx, db, None


def gelu(x, bias=None, approximate=False):
    return _GeluFn.apply(x, bias, approximate)


# =============================================================================================
# Vocab-parallel embedding lookup (csrc/embedding.cu): gather with the vocab-range mask and the [s, b, h] transpose folded
# in; backward scatter-adds straight into the fp32 main_grad of the table
# =============================================================================================

def _embedding_ref(ids, weight, vocab_start, sbh):
    local = ids - vocab_start
    mask = (local < 0) | (local >= weight.size(0))
    out = F.embedding(local.masked_fill(mask, 0), weight)
    out = out.masked_fill(mask.unsqueeze(-1), 0.0)
    return out.transpose(0, 1).contiguous() if sbh else out


def _embedding_kernel_ok(ids, weight):
    return (cuda_ops_available(weight) and ids.is_cuda and ids.dim() == 2 and weight.dim() == 2
            and weight.is_contiguous() and weight.size(1) % 8 == 0
            and weight.dtype in (torch.bfloat16, torch.float16, torch.float32))


class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, weight, vocab_start, sbh, accumulate_into_main_grad):
        idc = ids.contiguous().long()
        b, s = idc.shape
        out = torch.empty((s, b, weight.size(1)) if sbh else (b, s, weight.size(1)), dtype=weight.dtype,
                          device=weight.device)
        _C().embedding_fwd(idc, weight, out, vocab_start, sbh)
        _count()
        ctx.save_for_backward(idc)
        ctx.weight, ctx.vocab_start, ctx.sbh = weight, vocab_start, sbh
        ctx.fused = accumulate_into_main_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        (idc,) = ctx.saved_tensors
        w = ctx.weight
        dout = dout.contiguous()
        main_grad = getattr(w, "main_grad", None)
        if ctx.fused and main_grad is not None and main_grad.dtype == torch.float32 and main_grad.is_contiguous():
            _C().embedding_bwd(idc, dout, main_grad, ctx.vocab_start, ctx.sbh)
            _count()
            cb = getattr(w, "_grad_ready_callback", None)
            if cb is not None:
                cb()
            return None, None, None, None, None
        dw = torch.zeros(w.shape, dtype=torch.float32, device=w.device)
        _C().embedding_bwd(idc, dout, dw, ctx.vocab_start, ctx.sbh)
        _count()
        return None, dw.to(w.dtype), None, None, None


def embedding_lookup(ids, weight, vocab_start: int = 0, sbh: bool = False, accumulate_into_main_grad: bool = False):
    """ids [b, s] -> weight[ids - vocab_start] as [b, s, h] (or [s, b, h] with ``sbh``); ids outside this rank's vocab
    range give zero rows.  With ``accumulate_into_main_grad`` the backward adds into ``weight.main_grad`` (fp32)."""
    if _embedding_kernel_ok(ids, weight):
        return _EmbeddingFn.apply(ids, weight, int(vocab_start), bool(sbh), bool(accumulate_into_main_grad))
    return _embedding_ref(ids, weight, vocab_start, sbh)


# =============================================================================================
# bias + dropout + residual add (one kernel each way; the mask is a counter-based hash of (seed, element index), so
# nothing is stored for backward -- reference: megatron/model/transformer.py:563-609 jit-scripted bias_dropout_add)
# =============================================================================================

def _dropout_seed(n_elems: int) -> int:
    """63-bit seed drawn from (and advancing) the current CUDA generator, on the host: follows the model-parallel RNG
    tracker's forks and is reproduced exactly when activation recompute restores the generator state."""
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    seed, off = gen.initial_seed(), gen.get_offset()
    gen.set_offset(off + 4 * ((n_elems + 3) // 4))
    return (seed * 6364136223846793005 + off * 1442695040888963407 + 1) & 0x7FFFFFFFFFFFFFFF


def _bda_kernel_ok(x, bias, residual):
    F_ = x.size(-1)
    return (cuda_ops_available(x) and x.dtype in (torch.bfloat16, torch.float16, torch.float32) and F_ % 8 == 0
            and residual.dtype == x.dtype and residual.shape == x.shape
            and (bias is None or (bias.dtype == x.dtype and bias.numel() == F_)))


class _BiasDropoutAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, residual, p, seed):
        xc, rc = x.contiguous(), residual.contiguous()
        out = torch.empty_like(xc)
        _C().bias_dropout_add(xc, bias.contiguous() if bias is not None else None, rc, out, p, seed, False)
        _count()
        ctx.p, ctx.seed, ctx.has_bias = p, seed, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        if ctx.p > 0.0:
            dx = torch.empty_like(dout)
            _C().bias_dropout_add(dout, None, None, dx, ctx.p, ctx.seed, True)      # dx = mask * dout / (1 - p)
            _count()
        else:
            dx = dout
        dbias = dx.reshape(-1, dx.size(-1)).sum(0) if ctx.has_bias else None
        return dx, dbias, dout, None, None


def bias_dropout_add(x, bias, residual, p: float, training: bool):
    """residual + dropout(x + bias, p)."""
    p = float(p) if training else 0.0
    if _bda_kernel_ok(x, bias, residual):
        seed = _dropout_seed(x.numel()) if p > 0.0 else 0
        return _BiasDropoutAddFn.apply(x, bias, residual, p, seed)
    if bias is not None:
        x = x + bias
    if p > 0.0:
        x = F.dropout(x, p=p, training=True)
    return residual + x


# =============================================================================================
# Cross entropy statistics (used by parallel/cross_entropy.py)
# =============================================================================================

def ce_local_stats(logits2d, target, vocab_start):
    stats = torch.empty((logits2d.size(0), 4), dtype=torch.float32, device=logits2d.device)
    if logits2d.stride(1) != 1 or logits2d.stride(0) % 8 != 0:
        logits2d = logits2d.contiguous()
    _C().ce_stats(logits2d, target.contiguous(), stats, vocab_start)
    _count()
    return stats


def ce_backward_(logits2d, target, M, logS, g, vocab_start, smoothing, vocab_size):
    """Overwrites ``logits2d`` with the gradient and returns it."""
    _C().ce_bwd(logits2d, logits2d, target.contiguous(), M.contiguous(), logS.contiguous(), g, vocab_start,
                smoothing, vocab_size)
    _count()
    return logits2d


# =============================================================================================
# Softmax family (non-flash attention path)
# =============================================================================================

class _ScaledSoftmaxFn(torch.autograd.Function):
    """mode 0: softmax(scale*x); 1: with uint8 mask [b|1,1,sq,sk] (1 = masked); 2: causal."""

    @staticmethod
    def forward(ctx, x, mask, scale, mode):
        xc = x.contiguous()
        b, np_, sq, sk = xc.shape
        ctx.scale = scale
        if cuda_ops_available(xc) and xc.dtype in (torch.bfloat16, torch.float16, torch.float32):
            y = torch.empty_like(xc)
            m8 = mask.to(torch.uint8).contiguous() if mask is not None else None
            _C().softmax_fwd(xc, y, m8, scale, sq, sk, np_, mode)
            _count()
        else:
            z = xc.float() * scale
            if mode == 1:
                z = z.masked_fill(mask.bool(), -10000.0)
            elif mode == 2:
                causal = torch.ones(sq, sk, dtype=torch.bool, device=xc.device).tril(diagonal=sk - sq)
                z = z.masked_fill(~causal, float("-inf"))
            y = torch.softmax(z, dim=-1)
            if mode == 1:
                y = y * (~mask.bool().all(dim=-1, keepdim=True)).float()
            y = y.to(xc.dtype)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        if cuda_ops_available(y) and y.dtype in (torch.bfloat16, torch.float16, torch.float32):
            dx = dy.contiguous().clone()
            _C().softmax_bwd(dx, y, ctx.scale, y.size(-1))
            _count()
        else:
            yf, gf = y.float(), dy.float()
            dx = (ctx.scale * (yf * gf - yf * (yf * gf).sum(-1, keepdim=True))).to(y.dtype)
        return dx, None, None, None


def scaled_softmax(x, scale=1.0):
    return _ScaledSoftmaxFn.apply(x, None, scale, 0)


def scaled_masked_softmax(x, mask, scale=1.0):
    return _ScaledSoftmaxFn.apply(x, mask, scale, 1)


def scaled_upper_triang_masked_softmax(x, scale=1.0):
    """x [b, np, sq, sk] (or [attn_batches, sq, sk])."""
    if x.dim() == 3:
        return _ScaledSoftmaxFn.apply(x.unsqueeze(0), None, scale, 2).squeeze(0)
    return _ScaledSoftmaxFn.apply(x, None, scale, 2)


# =============================================================================================
# Flat-buffer optimizer primitives
# =============================================================================================

def accumulate_(dst_fp32: torch.Tensor, src: torch.Tensor) -> None:
    """dst(fp32) += src (any float dtype)."""
    if cuda_ops_available(dst_fp32) and dst_fp32.dtype == torch.float32 and src.is_contiguous() and \
            dst_fp32.is_contiguous() and src.dtype in (torch.bfloat16, torch.float16, torch.float32) and \
            src.data_ptr() % 16 == 0 and dst_fp32.data_ptr() % 16 == 0:
        _C().accumulate(src, dst_fp32)
        _count()
    else:
        dst_fp32.add_(src)


from .attention import flash_attention  # noqa: E402,F401
