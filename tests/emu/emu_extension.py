"""The attention entry points of the compiled extension (csrc/attention_bind.cpp: attn_fwd, attn_bwd, attn_fwd_packed,
attn_bwd_packed, attn_decode) backed by the kernel sources running on the CPU models of tests/emu -- a drop-in for
``ops._ext.load()`` in tests that want to drive the real kernels without a GPU (CPU tensors, same argument lists, same
output layouts as the bindings)."""
import ctypes
import os

import torch

import host_build


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class EmuExtension:
    def __init__(self, build_dir):
        self.attn = ctypes.CDLL(host_build.build(["attention_sm100.cu", "attention_bwd_sm100.cu"], os.path.join(build_dir, "attn")))
        self.dec = ctypes.CDLL(host_build.build(["attention_decode.cu"], os.path.join(build_dir, "dec")))

    # ---- separate q / k / v: [b, s, n, hn] views, hn contiguous
    @staticmethod
    def _str(t):
        return (ctypes.c_longlong * 3)(t.stride(2), t.stride(1), t.stride(0))

    def attn_fwd(self, q, k, v, causal, window, scale, dropout_p=0.0, seed=0):
        assert causal
        b, s, n, hn = q.shape
        nkv = k.size(2)
        out = torch.empty(s, b, n, hn, dtype=q.dtype)
        lse = torch.empty(b, n, s)
        hm = (ctypes.c_int * 6)(n // nkv, 0, 1, 0, 1, 0)
        rc = self.attn.mlb_attn_fwd_ex(_p(q), _p(k), _p(v), self._str(q), self._str(k), self._str(v), n, nkv, nkv, hm, n // nkv,
                                       s, b, n, int(window), ctypes.c_float(scale), _p(out), ctypes.c_longlong(b * n * hn),
                                       ctypes.c_longlong(n * hn), _p(lse), hn, int(q.dtype == torch.float16),
                                       ctypes.c_float(dropout_p), ctypes.c_ulonglong(seed), None)
        assert rc == 0, rc
        return [out.permute(1, 0, 2, 3), lse]

    def attn_bwd(self, dout, q, k, v, out, lse, causal, window, scale, dropout_p=0.0, seed=0):
        b, s, n, hn = q.shape
        nkv = k.size(2)
        dq, dk, dv = (torch.empty(s, b, h, hn, dtype=q.dtype).permute(1, 0, 2, 3) for h in (n, nkv, nkv))
        delta = torch.empty(b, n, s)
        hm = (ctypes.c_int * 6)(n // nkv, 0, 1, 0, 1, 0)
        S = self._str
        rc = self.attn.mlb_attn_bwd_ex(_p(q), _p(k), _p(v), _p(out), _p(dout), S(q), S(k), S(v), S(out), S(dout), n, nkv, nkv, hm,
                                       n // nkv, s, b, n, int(window), ctypes.c_float(scale), _p(lse), _p(delta), _p(dq), _p(dk),
                                       _p(dv), S(dq), S(dk), S(dv), hn, int(q.dtype == torch.float16), ctypes.c_float(dropout_p),
                                       ctypes.c_ulonglong(seed), None)
        assert rc == 0, rc
        return [dq, dk, dv]

    # ---- packed QKV [s, b, nkv * (g + 2) * hn]
    def attn_fwd_packed(self, mixed, nkv, g, window, scale, hn, dropout_p=0.0, seed=0):
        s, b = mixed.shape[:2]
        n, mh = nkv * g, nkv * (g + 2)
        ms = (ctypes.c_longlong * 3)(hn, mixed.stride(0), mixed.stride(1))
        out = torch.empty(s, b, n * hn, dtype=mixed.dtype)
        lse = torch.empty(b, n, s)
        hm = (ctypes.c_int * 6)(g + 2, 0, g + 2, g, g + 2, g + 1)
        rc = self.attn.mlb_attn_fwd_ex(_p(mixed), _p(mixed), _p(mixed), ms, ms, ms, mh, mh, mh, hm, g, s, b, n, int(window),
                                       ctypes.c_float(scale), _p(out), ctypes.c_longlong(b * n * hn), ctypes.c_longlong(n * hn),
                                       _p(lse), hn, int(mixed.dtype == torch.float16), ctypes.c_float(dropout_p),
                                       ctypes.c_ulonglong(seed), None)
        assert rc == 0, rc
        return [out, lse]

    def attn_bwd_packed(self, dout, mixed, out, lse, nkv, g, window, scale, hn, dropout_p=0.0, seed=0):
        s, b = mixed.shape[:2]
        n, mh = nkv * g, nkv * (g + 2)
        L3 = lambda t: (ctypes.c_longlong * 3)(hn, t.stride(0), t.stride(1))
        dmixed = torch.empty_like(mixed)
        delta = torch.empty(b, n, s)
        hm = (ctypes.c_int * 6)(g + 2, 0, g + 2, g, g + 2, g + 1)
        rc = self.attn.mlb_attn_bwd_ex(_p(mixed), _p(mixed), _p(mixed), _p(out), _p(dout), L3(mixed), L3(mixed), L3(mixed), L3(out),
                                       L3(dout), mh, mh, mh, hm, g, s, b, n, int(window), ctypes.c_float(scale), _p(lse), _p(delta),
                                       _p(dmixed), _p(dmixed), _p(dmixed), L3(dmixed), L3(dmixed), L3(dmixed), hn,
                                       int(mixed.dtype == torch.float16), ctypes.c_float(dropout_p), ctypes.c_ulonglong(seed), None)
        assert rc == 0, rc
        return dmixed

    # ---- KV-cache decode (same split heuristic as the binding)
    def attn_decode(self, q, k, v, window, scale, splits):
        b, sq, n, hn = q.shape
        sk, nkv = k.size(1), k.size(2)
        n_splits = splits if splits > 0 else max(1, min((296 + b * nkv - 1) // (b * nkv), (sk + 255) // 256))
        kps = (((sk + n_splits - 1) // n_splits) + 31) // 32 * 32
        n_splits = (sk + kps - 1) // kps
        rows = b * nkv * n_splits * sq * (n // nkv)
        part_o, part_ml = torch.empty(rows, hn), torch.empty(rows, 2)
        out = torch.empty(b, sq, n, hn, dtype=q.dtype)
        st = lambda t: (ctypes.c_longlong * 3)(t.stride(0), t.stride(1), t.stride(2))
        rc = self.dec.mlb_attn_decode(0 if q.dtype == torch.bfloat16 else 1, _p(q), _p(k), _p(v), st(q), st(k), st(v), b, sq, sk,
                                      n, nkv, hn, int(window), ctypes.c_float(scale), n_splits, kps, _p(part_o), _p(part_ml),
                                      _p(out), None)
        assert rc == 0, rc
        return out
