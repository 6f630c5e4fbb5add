"""The tcgen05 attention kernels (csrc/attention_sm100.cu, attention_bwd_sm100.cu: forward single-tile and two-tile,
delta, dK/dV, dQ) executed on the CPU.  Their real source is compiled for the host against a FUNCTIONAL model of what
they are written on (tests/emu/cuda_emu/tcgen05_model.h, ptx.cuh): mbarrier phases and transaction counts, TMA tile
loads through the 128-byte swizzle, tensor memory, tcgen05.mma reading its operands through the shared-memory /
instruction descriptors the real code builds, tcgen05.ld / st, named barriers -- one OS thread per CUDA thread.

The bf16 / no-dropout kernels are validated on B200 (tests/test_ops_gpu.py); that they also pass here pins the model.
On that footing the variants that have not run on hardware yet (fp16 operands, attention dropout with the mask
regenerated in three tile layouts) are checked against the fp32 oracle, and ThreadSanitizer checks the barrier
protocol (it found a phase-aliasing race in the single-tile forward kernel: see ``o_done`` in attention_sm100.cu)."""
import ctypes
import math
import os
import subprocess
import sys

import pytest
import torch

from megatron_llm_b200.ops.attention import attention_reference, dropout_keep_mask

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import host_build  # noqa: E402

FILES = ["attention_sm100.cu", "attention_bwd_sm100.cu"]
SEED = 0x1357_9BDF_0246_8ACE


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    return ctypes.CDLL(host_build.build(FILES, str(tmp_path_factory.mktemp("emu_attn"))))


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _str(t):     # [b, s, n, hn] view -> (head, seq, batch) strides in elements
    return (ctypes.c_longlong * 3)(t.stride(2), t.stride(1), t.stride(0))


def _fwd_bwd(lib, q, k, v, do, window, p, head_map=None, map_heads=None):
    b, s, n, hn = q.shape
    nkv = k.size(2)
    g = n // nkv
    w = -1 if window is None else window
    hm = (ctypes.c_int * 6)(*(head_map or (g, 0, 1, 0, 1, 0)))
    mq, mk = map_heads or (n, nkv)
    fp16, scale, seed = int(q.dtype == torch.float16), ctypes.c_float(1.0 / math.sqrt(hn)), ctypes.c_ulonglong(SEED if p else 0)
    out = torch.full((s, b, n, hn), float("nan"), dtype=q.dtype).permute(1, 0, 2, 3)
    lse, delta = torch.full((b, n, s), float("nan")), torch.zeros(b, n, s)
    assert lib.mlb_attn_fwd_ex(_p(q), _p(k), _p(v), _str(q), _str(k), _str(v), mq, mk, mk, hm, g, s, b, n, w, scale, _p(out),
                               ctypes.c_longlong(b * n * hn), ctypes.c_longlong(n * hn), _p(lse), hn, fp16,
                               ctypes.c_float(p), seed, None) == 0
    dq, dk, dv = (torch.full((s, b, h, hn), float("nan"), dtype=q.dtype).permute(1, 0, 2, 3) for h in (n, nkv, nkv))
    assert lib.mlb_attn_bwd_ex(_p(q), _p(k), _p(v), _p(out), _p(do), _str(q), _str(k), _str(v), _str(out), _str(do), mq, mk,
                               mk, hm, g, s, b, n, w, scale, _p(lse), _p(delta), _p(dq), _p(dk), _p(dv), _str(dq),
                               _str(dk), _str(dv), hn, fp16, ctypes.c_float(p), seed, None) == 0
    return out, lse, dq, dk, dv


def _rel(a, r):
    return ((a.float() - r).norm() / r.norm()).item()


@pytest.mark.parametrize("dtype,b,s,n,nkv,hn,window,p", [
    (torch.bfloat16, 1, 256, 2, 1, 128, None, 0.0),      # two-tile forward kernel, GQA
    (torch.bfloat16, 1, 384, 2, 2, 128, 200, 0.0),       # single-tile forward kernel (3 KV tiles), sliding window
    (torch.bfloat16, 2, 256, 4, 1, 64, None, 0.0),       # head dim 64, MQA, batch 2
    (torch.float16, 1, 256, 2, 1, 128, None, 0.0),       # fp16 operands
    (torch.float16, 1, 384, 2, 2, 64, 100, 0.0),
    (torch.bfloat16, 1, 256, 2, 1, 128, None, 0.1),      # dropout
    (torch.bfloat16, 1, 384, 2, 2, 64, 100, 0.25),
    (torch.bfloat16, 1, 640, 1, 1, 128, None, 0.1),      # 5 KV tiles through the single-tile kernel
    (torch.float16, 1, 256, 4, 2, 128, None, 0.5),       # fp16 + dropout
])
def test_attention_kernels_on_the_functional_model(lib, dtype, b, s, n, nkv, hn, window, p):
    g = torch.Generator().manual_seed(s + n)
    q, k, v, do = (torch.randn(b, s, h, hn, generator=g).to(dtype) for h in (n, nkv, nkv, n))
    out, lse, dq, dk, dv = _fwd_bwd(lib, q, k, v, do, window, p)
    qf, kf, vf = (x.float().requires_grad_() for x in (q, k, v))
    keep = dropout_keep_mask(SEED, p, b, n, s, s) if p > 0 else None
    ref = attention_reference(qf, kf, vf, True, window, None, p, keep)
    ref.backward(do.float())
    tol = 6e-3 if dtype == torch.bfloat16 else 1e-3
    assert _rel(out, ref) < tol and _rel(dq, qf.grad) < tol and _rel(dk, kf.grad) < tol and _rel(dv, vf.grad) < tol
    sc = torch.einsum("bqnh,bknh->bnqk", qf.detach(), kf.detach().repeat_interleave(n // nkv, dim=2)) / math.sqrt(hn)
    qi, ki = torch.arange(s)[:, None], torch.arange(s)[None, :]
    allowed = (ki <= qi) & ((ki >= qi - window) if window is not None else True)
    assert (lse - torch.logsumexp(sc.masked_fill(~allowed, float("-inf")), -1)).abs().max().item() < 1e-3


def test_packed_qkv_addressing_on_the_functional_model(lib):
    """Q / K / V read in place from the packed projection output [s, b, nkv * (g + 2) * hn] through the head map, and
    dQ / dK / dV written into one packed gradient buffer of the same layout."""
    torch.manual_seed(3)
    s, b, nkv, g, hn = 256, 2, 2, 2, 64
    n = nkv * g
    mixed = torch.randn(s, b, nkv * (g + 2) * hn).bfloat16()
    do = torch.randn(s, b, n, hn).bfloat16().permute(1, 0, 2, 3)
    hm, mh = (g + 2, 0, g + 2, g, g + 2, g + 1), nkv * (g + 2)
    # every operand of the C entry points is the packed buffer
    w, scale = -1, ctypes.c_float(1.0 / math.sqrt(hn))
    hmc = (ctypes.c_int * 6)(*hm)
    ms = (ctypes.c_longlong * 3)(hn, mixed.stride(0), mixed.stride(1))
    out = torch.full((s, b, n, hn), float("nan")).bfloat16()
    os_ = (ctypes.c_longlong * 3)(hn, out.stride(0), out.stride(1))
    lse, delta = torch.zeros(b, n, s), torch.zeros(b, n, s)
    assert lib.mlb_attn_fwd_ex(_p(mixed), _p(mixed), _p(mixed), ms, ms, ms, mh, mh, mh, hmc, g, s, b, n, w, scale, _p(out),
                               ctypes.c_longlong(b * n * hn), ctypes.c_longlong(n * hn), _p(lse), hn, 0, ctypes.c_float(0),
                               ctypes.c_ulonglong(0), None) == 0
    dmixed = torch.full_like(mixed.float(), float("nan")).bfloat16()
    doc = do.permute(1, 0, 2, 3).contiguous()                                    # [s, b, n, hn]
    ds = (ctypes.c_longlong * 3)(hn, doc.stride(0), doc.stride(1))
    assert lib.mlb_attn_bwd_ex(_p(mixed), _p(mixed), _p(mixed), _p(out), _p(doc), ms, ms, ms, os_, ds, mh, mh, mh, hmc, g, s,
                               b, n, w, scale, _p(lse), _p(delta), _p(dmixed), _p(dmixed), _p(dmixed), ms, ms, ms, hn, 0,
                               ctypes.c_float(0), ctypes.c_ulonglong(0), None) == 0
    qkv = mixed.view(s, b, nkv, g + 2, hn)
    q = qkv[:, :, :, :g].reshape(s, b, n, hn).transpose(0, 1).float().requires_grad_()
    k = qkv[:, :, :, g].transpose(0, 1).float().requires_grad_()
    v = qkv[:, :, :, g + 1].transpose(0, 1).float().requires_grad_()
    ref = attention_reference(q, k, v, True)
    ref.backward(do.float())
    dm = dmixed.view(s, b, nkv, g + 2, hn)
    assert _rel(out.transpose(0, 1), ref) < 6e-3
    assert _rel(dm[:, :, :, :g].reshape(s, b, n, hn).transpose(0, 1), q.grad) < 6e-3
    assert _rel(dm[:, :, :, g].transpose(0, 1), k.grad) < 6e-3 and _rel(dm[:, :, :, g + 1].transpose(0, 1), v.grad) < 6e-3


def test_attention_barrier_protocol_under_thread_sanitizer(tmp_path):
    """All five kernels (dropout instantiation), single-tile and two-tile forward, head dims 128 and 64: no access to
    tensor / shared memory that is not ordered by the kernels' own barriers."""
    exe = host_build.build_race_driver(FILES, str(tmp_path), name="attn_race", driver="attn_race_driver.cpp")
    for args in (["384", "128"], ["512", "64"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (args, r.stderr[-3000:])


def test_python_attention_path_on_the_real_kernels(monkeypatch, tmp_path_factory):
    """ops/attention_sm100.py driven end to end on CPU tensors with the REAL kernel sources behind the extension API
    (tests/emu/emu_extension.py): the first-use self-tests of the fp16 / dropout / decode variants pass with their
    production thresholds, and ``attention()`` -- zero-padding of a ragged length, dropout seed plumbing, autograd --
    reproduces the fp32 oracle with the same mask."""
    import emu_extension
    from megatron_llm_b200.ops import _ext, attention_sm100
    ext = emu_extension.EmuExtension(str(tmp_path_factory.mktemp("emu_ext")))
    monkeypatch.setattr(_ext, "load", lambda: ext)
    monkeypatch.setattr(attention_sm100, "_feature_state", {})
    for var in ("MLB200_ATTN", "MLB200_DISABLE_KERNELS", "MLB200_ATTN_FP16", "MLB200_ATTN_DROPOUT", "MLB200_ATTN_DECODE"):
        monkeypatch.delenv(var, raising=False)
    for var in [k for k in os.environ if k.startswith("MLB200_ATTN_SELFTEST_")]:
        os.environ.pop(var)      # verdicts inherited from whatever ran earlier in this process say nothing about these kernels
    dev = torch.device("cpu")
    assert attention_sm100.feature_ok("dropout", 128, torch.bfloat16, dev)
    assert attention_sm100.feature_ok("fp16", 64, torch.float16, dev)
    assert attention_sm100.feature_ok("decode", 128, torch.bfloat16, dev)
    for var in [k for k in os.environ if k.startswith("MLB200_ATTN_SELFTEST_")]:
        os.environ.pop(var)      # verdicts about the emulated kernels must not reach processes started later
    monkeypatch.setattr(attention_sm100, "_draw_seed", lambda p, n: SEED if p > 0 else 0)
    torch.manual_seed(5)
    b, s, n, nkv, hn, window, p = 1, 200, 4, 2, 128, 150, 0.1
    q, k, v = (torch.randn(b, s, h, hn).bfloat16().requires_grad_() for h in (n, nkv, nkv))
    do = torch.randn(b, s, n, hn).bfloat16()
    assert attention_sm100.supported(q, k, v, True, window, p)
    out = attention_sm100.attention(q, k, v, True, window, None, p)
    out.backward(do)
    qf, kf, vf = (t.detach().float().requires_grad_() for t in (q, k, v))
    ref = attention_reference(qf, kf, vf, True, window, None, p, dropout_keep_mask(SEED, p, b, n, s, s))
    ref.backward(do.float())
    assert _rel(out, ref) < 6e-3 and _rel(q.grad, qf.grad) < 6e-3 and _rel(k.grad, kf.grad) < 6e-3 and _rel(v.grad, vf.grad) < 6e-3
    with torch.no_grad():                                       # the decode step through decode_supported / decode_attention
        kc, vc = (torch.randn(97, b, nkv, hn).bfloat16() for _ in range(2))
        q1 = torch.randn(b, 1, n, hn).bfloat16()
        kk, vv = kc.transpose(0, 1), vc.transpose(0, 1)
        assert attention_sm100.decode_supported(q1, kk, vv, True, 0.0)
        o1 = attention_sm100.decode_attention(q1, kk, vv, None, None)
        assert _rel(o1, attention_reference(q1.float(), kk.float(), vv.float(), True)) < 6e-3


def test_hardware_check_scripts_run_on_the_model(tmp_path_factory):
    """The scripts of tests/test_z_attention_variants_gpu.py (the first hardware run of the new attention variants) are
    themselves executed here, on CPU tensors against the emulated extension, so that a mistake in a script cannot be
    mistaken for a kernel failure on the device.  One script runs in the suite; all of the direct-kernel ones
    (fp16_training, dropout_training, packed_dropout, decode, single_tile_forward_kernel) passed this way when they
    were written (about six minutes)."""
    import importlib.util
    import emu_extension
    spec = importlib.util.spec_from_file_location("tz", os.path.join(ROOT, "tests", "test_z_attention_variants_gpu.py"))
    tz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tz)
    ext = emu_extension.EmuExtension(str(tmp_path_factory.mktemp("emu_ext_z")))
    prelude = (tz.PRELUDE % {"root": tz.ROOT}).replace("mod = _ext.load()", "mod = _EMU_EXT") \
        .replace('dev = torch.device("cuda:0")', 'dev = torch.device("cpu")')
    assert "_EMU_EXT" in prelude and '"cpu"' in prelude
    for name in os.environ.get("MLB200_ZCHECKS_ON_MODEL", "packed_dropout").split(","):
        env = {"_EMU_EXT": ext}
        exec(compile(prelude + tz.CHECKS[name], name, "exec"), env)
        bad = {f"{c}.{k}": v for c, d in env["res"].items() for k, v in d.items() if not (v == v and v < 2e-2)}
        assert not bad, (name, bad)
