"""Python side of the hand-written sm_100a attention kernels (csrc/attention_sm100.cu, attention_bwd_sm100.cu,
attention_decode.cu).

Envelope: causal self-attention, head_dim 128 or 64, bf16 or fp16, GQA / MQA, sliding window, attention dropout
(counter-based mask, csrc/attention_dropout.cuh).  sq == sk (training, and the prompt pass of text generation) runs on
the tcgen05 kernels, with lengths that are not a multiple of the 128-row tile zero-padded; the KV-cache decode step
(a few query positions against a long cache) runs on the split-KV kernel.

Kernel variants that could not be run on hardware before they were merged (``_FEATURES``) are guarded by a one-time
numerical self-test against the fp32 oracle on first use: a variant that fails it is reported loudly and the call falls
back to the ``flash_attn`` library instead of producing wrong numbers.  On a GPU the self-test runs in a throw-away
child process with a time limit (``_selftest_isolated``), so that a faulting or dead-locked first launch cannot poison
the CUDA context of the job; the verdict is handed to the processes started afterwards through the environment.
``MLB200_ATTN_<FEATURE>=1`` skips the self-test, ``=0`` disables the variant."""
from __future__ import annotations

import math
import os
import sys
import warnings

import torch
import torch.nn.functional as F

from . import _ext

_FEATURES = {"fp16": "MLB200_ATTN_FP16", "dropout": "MLB200_ATTN_DROPOUT", "decode": "MLB200_ATTN_DECODE"}
_feature_state = {}      # (feature, head_dim) -> bool


def _kernels_enabled() -> bool:
    return not (os.environ.get("MLB200_ATTN", "1") == "0" or os.environ.get("MLB200_DISABLE_KERNELS", "0") == "1")


def _head_dim_ok(hn: int) -> bool:
    """128 (Llama / Mistral) and 64 (Falcon-40B: 128 heads x 64, GPT-2 style models); both instantiations are
    validated against the fp32 reference in tests/test_ops_gpu.py (``MLB200_ATTN_HD64=0`` routes 64 to the library)."""
    return hn == 128 or (hn == 64 and os.environ.get("MLB200_ATTN_HD64", "1") == "1")


def _rel_err(a, b) -> float:
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _selftest_training(hn: int, dtype, dropout_p: float, device) -> float:
    """forward + dq / dk / dv of the tcgen05 kernels vs the fp32 oracle on a small GQA problem (worst relative error)."""
    from .attention import attention_reference, dropout_keep_mask
    mod = _ext.load()
    g = torch.Generator(device=device).manual_seed(20240607)      # never touches the global RNG streams
    b, s, nq, nkv = 1, 256, 4, 2
    q, k, v, do = (torch.randn(b, s, n, hn, device=device, generator=g).to(dtype) for n in (nq, nkv, nkv, nq))
    seed = 0x5EED_1234_ABCD_0001 if dropout_p > 0 else 0
    scale = 1.0 / math.sqrt(hn)
    out, lse = mod.attn_fwd(q, k, v, True, -1, scale, dropout_p, seed)
    dq, dk, dv = mod.attn_bwd(do, q, k, v, out, lse, True, -1, scale, dropout_p, seed)
    with torch.enable_grad():
        qf, kf, vf = (t.float().requires_grad_() for t in (q, k, v))
        keep = dropout_keep_mask(seed, dropout_p, b, nq, s, s, device=device) if dropout_p > 0 else None
        ref = attention_reference(qf, kf, vf, True, None, scale, dropout_p, keep)
        ref.backward(do.float())
    return max(_rel_err(out, ref), _rel_err(dq, qf.grad), _rel_err(dk, kf.grad), _rel_err(dv, vf.grad))


def _selftest_decode(hn: int, dtype, device) -> float:
    from .attention import attention_reference
    mod = _ext.load()
    g = torch.Generator(device=device).manual_seed(20240608)
    worst = 0.0
    for b, sq, sk, nq, nkv, window in ((2, 1, 333, 8, 2, None), (1, 2, 200, 4, 4, 64)):
        q = torch.randn(b, sq, nq, hn, device=device, generator=g).to(dtype)
        kmem = torch.randn(sk + 3, b + 1, nkv, hn, device=device, generator=g).to(dtype)
        vmem = torch.randn(sk + 3, b + 1, nkv, hn, device=device, generator=g).to(dtype)
        k, v = kmem[:sk, 1:].transpose(0, 1), vmem[:sk, 1:].transpose(0, 1)
        out = mod.attn_decode(q, k, v, -1 if window is None else window, 1.0 / math.sqrt(hn), 0)
        worst = max(worst, _rel_err(out, attention_reference(q.float(), k.float(), v.float(), True, window)))
    return worst


def _selftest(feature: str, hn: int, device) -> float:
    with torch.no_grad():
        if feature == "decode":
            return max(_selftest_decode(hn, torch.bfloat16, device), _selftest_decode(hn, torch.float16, device))
        if feature == "fp16":
            return _selftest_training(hn, torch.float16, 0.0, device)
        return _selftest_training(hn, torch.bfloat16, 0.1, device)


def _verdict_env(feature: str, hn: int) -> str:
    return f"MLB200_ATTN_SELFTEST_{feature.upper()}_HD{hn}"


def _selftest_isolated(feature: str, hn: int, device) -> float:
    """The self-test in a short-lived child process with a time limit.  A first-ever launch of a kernel variant that
    faults leaves a sticky error in its CUDA context and one that dead-locks never returns; neither may take the
    training job (or a test session) with it, so the context that finds out is a throw-away one."""
    import subprocess
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    env["MLB200_ATTN_SELFTEST_INPROC"] = "1"
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    limit = float(os.environ.get("MLB200_ATTN_SELFTEST_TIMEOUT", "180"))
    try:
        r = subprocess.run([sys.executable, "-m", "megatron_llm_b200.ops.attention_sm100", feature, str(hn), str(index)],
                           env=env, capture_output=True, text=True, timeout=limit)
    except subprocess.TimeoutExpired:
        raise RuntimeError(f"self-test did not finish within {limit:.0f} s (killed)") from None
    for line in reversed(r.stdout.splitlines()):
        if line.startswith("MLB200_SELFTEST_ERR "):
            return float(line.split()[1])
    tail = (r.stderr or r.stdout).strip().splitlines()[-3:]
    raise RuntimeError(f"self-test process exited with code {r.returncode}: " + " | ".join(tail))


def feature_ok(feature: str, hn: int, dtype, device) -> bool:
    """May this kernel variant be used?  Decided once per process and (variant, head_dim) -- see the module docstring."""
    key = (feature, hn)
    if key in _feature_state:
        return _feature_state[key]
    mode = os.environ.get(_FEATURES[feature], "auto")
    if mode in ("0", "1"):
        _feature_state[key] = mode == "1"
        return _feature_state[key]
    if torch.device(device).type == "cuda" and torch.cuda.is_current_stream_capturing():
        return False                        # cannot self-test inside a graph capture; decided on the next eager call
    inherited = os.environ.get(_verdict_env(feature, hn))      # a parent process of this job has already decided
    if inherited in ("0", "1"):
        _feature_state[key] = inherited == "1"
        return _feature_state[key]
    try:
        if torch.device(device).type == "cuda" and os.environ.get("MLB200_ATTN_SELFTEST_INPROC", "0") != "1":
            err = _selftest_isolated(feature, hn, device)
        else:
            err = _selftest(feature, hn, device)
        ok = err == err and err < 3e-2
        detail = f"relative error {err:.3e}"
    except Exception as e:                  # a launch / binding failure is a failed self-test, not a crash
        ok, detail = False, f"{type(e).__name__}: {e}"
    _feature_state[key] = ok
    os.environ[_verdict_env(feature, hn)] = "1" if ok else "0"      # ranks / tools started from here inherit it
    if not ok:
        msg = (f"megatron_llm_b200: the sm_100a attention variant '{feature}' (head_dim {hn}) FAILED its self-test "
               f"({detail}); falling back to the flash_attn library for it")
        warnings.warn(msg)
        print(msg, file=sys.stderr, flush=True)
    return ok


def _variants_ok(t, hn: int, dropout_p: float) -> bool:
    if t.dtype == torch.float16 and not feature_ok("fp16", hn, t.dtype, t.device):
        return False
    if dropout_p > 0.0 and not feature_ok("dropout", hn, t.dtype, t.device):
        return False
    # fp16 + dropout is the combination of two separately tested code paths (the flags are independent)
    return t.dtype in (torch.bfloat16, torch.float16) and 0.0 <= dropout_p < 1.0


def supported(q, k, v, causal, window, dropout_p) -> bool:
    """Self-attention (sq == sk) on the tcgen05 kernels, forward + backward.  Sequence lengths that are not a multiple
    of the 128-row tile (variable-length instruction tuning) are zero-padded by ``attention``."""
    if not _kernels_enabled():
        return False
    _ext.load()      # a CUDA tensor without the built extension is an error, never a silent library fallback
    hn = q.size(-1)
    return (_head_dim_ok(hn) and causal and q.size(1) == k.size(1) and q.size(2) % k.size(2) == 0
            and _variants_ok(q, hn, dropout_p))


def _draw_seed(dropout_p: float, n_elems: int) -> int:
    if dropout_p <= 0.0:
        return 0
    from . import _dropout_seed          # advances the current CUDA generator (follows the TP RNG tracker / recompute)
    return _dropout_seed(n_elems)


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal, window, scale, dropout_p, seed):
        mod = _ext.load()
        out, lse = mod.attn_fwd(q, k, v, causal, -1 if window is None else int(window), float(scale), float(dropout_p),
                                int(seed))
        _ext.count()
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.causal, ctx.window, ctx.scale, ctx.dropout = causal, window, scale, (float(dropout_p), int(seed))
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        mod = _ext.load()
        dq, dk, dv = mod.attn_bwd(dout if dout.stride(-1) == 1 else dout.contiguous(), q, k, v, out, lse, ctx.causal,
                                  -1 if ctx.window is None else int(ctx.window), float(ctx.scale), *ctx.dropout)
        _ext.count(3)
        return dq, dk, dv, None, None, None, None, None


def attention(q, k, v, causal, window, scale, dropout_p: float = 0.0):
    scale = scale if scale is not None else 1.0 / math.sqrt(q.size(-1))
    seed = _draw_seed(dropout_p, q.size(0) * q.size(2) * q.size(1) * k.size(1))
    sq = q.size(1)
    pad = (-sq) % 128
    if pad:
        # zero rows up to the tile: padded keys lie in every real query's future (causal), padded query rows are
        # dropped again and receive a zero output gradient, so they contribute nothing to dK / dV either; autograd
        # differentiates through pad / slice
        q, k, v = (F.pad(t, (0, 0, 0, 0, 0, pad)) for t in (q, k, v))
    out = _AttnFn.apply(q, k, v, causal, window, scale, dropout_p, seed)
    return out[:, :sq] if pad else out


# ------------------------------------------------------------------------------------------------ inference shapes
_DECODE_MAX_ROWS = 64        # query rows per KV group (sq * g) the split-KV kernel takes (8 per pass over the cache)


def _decode_aligned(*ts) -> bool:
    return all(t.stride(3) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and t.stride(2) % 8 == 0
               and t.data_ptr() % 16 == 0 for t in ts)


def decode_supported(q, k, v, causal, dropout_p) -> bool:
    """The KV-cache step of text generation: a few query positions against a longer cache, no gradient needed.
    (The prompt pass has sq == sk and runs on the tcgen05 forward kernel through ``supported`` / ``attention``.)"""
    if not _kernels_enabled() or not causal or dropout_p != 0.0 or q.dtype not in (torch.bfloat16, torch.float16):
        return False
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        return False
    hn, sq, sk = q.size(-1), q.size(1), k.size(1)
    if not _head_dim_ok(hn) or q.size(2) % k.size(2) != 0 or sk < sq:
        return False
    _ext.load()
    return (sq * (q.size(2) // k.size(2)) <= _DECODE_MAX_ROWS and _decode_aligned(q, k, v)
            and feature_ok("decode", hn, q.dtype, q.device))


def decode_attention(q, k, v, window, scale):
    """The new positions against the KV cache on the split-KV kernel (csrc/attention_decode.cu); the causal mask is
    aligned bottom-right (query i sits at position sk - sq + i)."""
    mod = _ext.load()
    scale = scale if scale is not None else 1.0 / math.sqrt(q.size(-1))
    out = mod.attn_decode(q, k, v, -1 if window is None else int(window), float(scale), 0)
    _ext.count(2)
    return out


# ------------------------------------------------------------------------------------------------ packed QKV path
def packed_supported(mixed, nkv, g, hn, dropout_p) -> bool:
    """``mixed`` = QKV projection output [s, b, nkv * (g + 2) * hn] (per KV group: g query heads, then k, then v)."""
    if not _kernels_enabled() or os.environ.get("MLB200_ATTN_PACKED", "1") == "0":
        return False
    if not (mixed.is_cuda and mixed.dtype in (torch.bfloat16, torch.float16) and _head_dim_ok(hn)
            and mixed.dim() == 3 and mixed.stride(2) == 1 and mixed.size(0) % 128 == 0):
        return False
    _ext.load()      # (fails loudly on a GPU box without the extension)
    return _variants_ok(mixed, hn, dropout_p)


class _PackedAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mixed, nkv, g, window, scale, hn, dropout_p, seed):
        mod = _ext.load()
        out, lse = mod.attn_fwd_packed(mixed, nkv, g, -1 if window is None else int(window), float(scale), hn,
                                       float(dropout_p), int(seed))
        _ext.count()
        ctx.save_for_backward(mixed, out, lse)
        ctx.cfg = (nkv, g, window, scale, hn)
        ctx.dropout = (float(dropout_p), int(seed))
        return out

    @staticmethod
    def backward(ctx, dout):
        mixed, out, lse = ctx.saved_tensors
        nkv, g, window, scale, hn = ctx.cfg
        mod = _ext.load()
        d = dout if dout.stride(-1) == 1 else dout.contiguous()
        dmixed = mod.attn_bwd_packed(d, mixed, out, lse, nkv, g, -1 if window is None else int(window), float(scale),
                                     hn, *ctx.dropout)
        _ext.count(3)
        return dmixed, None, None, None, None, None, None, None


def packed_attention(mixed, nkv, g, window=None, scale=None, hn=None, dropout_p: float = 0.0):
    """Causal attention straight from the packed (already rotated) QKV buffer -> context [s, b, nkv * g * hn]."""
    hn = hn if hn is not None else mixed.size(-1) // (nkv * (g + 2))
    scale = scale if scale is not None else 1.0 / math.sqrt(hn)
    seed = _draw_seed(dropout_p, mixed.size(1) * nkv * g * mixed.size(0) * mixed.size(0))
    return _PackedAttnFn.apply(mixed, nkv, g, window, scale, hn, dropout_p, seed)


if __name__ == "__main__":      # child of _selftest_isolated: <feature> <head_dim> <cuda device index>
    _feature, _hn, _index = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    torch.cuda.set_device(_index)
    _err = _selftest(_feature, _hn, torch.device("cuda", _index))
    torch.cuda.synchronize()
    print(f"MLB200_SELFTEST_ERR {_err:.6e}", flush=True)
