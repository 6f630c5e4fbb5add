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


DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def _o(t):
    return ctypes.c_void_p(None if t is None else t.data_ptr())


class FullEmuExtension(EmuExtension):
    """Every binding the training stack calls (csrc/bindings.cpp), on the CPU models: the tcgen05 GEMMs (1-CTA / 2-CTA),
    attention, and the SIMT kernels.  With ``ops._ext.load`` pointing here and the ``is_cuda`` predicates of
    ``ops/__init__.py`` lifted (see tests/test_model_on_emulated_kernels.py), a whole training step runs through the real
    kernel sources on CPU tensors."""

    def __init__(self, build_dir):
        super().__init__(build_dir)
        self.gm = ctypes.CDLL(host_build.build(["gemm_sm100.cu", "gemm2_sm100.cu"], os.path.join(build_dir, "gemm")))
        self.simt = ctypes.CDLL(host_build.build(["ce.cu", "softmax.cu", "norm.cu", "elementwise.cu", "optim.cu", "embedding.cu"],
                                                 os.path.join(build_dir, "simt")))
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def num_sms(self):
        return 4

    def gemm(self, A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, epilogue, block_n, comm, sms):
        assert comm is None
        self._count("gemm")
        fp16 = int(A.dtype == torch.float16)
        out_bytes = 2 if epilogue in (0, 3) else 4
        if block_n == 512 and (ldc * out_bytes) % 16 == 0:
            rc = self.gm.mlb_gemm_bf16_2cta(_p(A), _p(B), _p(C), M, N, K, lda, ldb, ldc, int(a_mn), int(b_mn), epilogue, fp16, 4, None)
        else:
            rc = self.gm.mlb_gemm_bf16(_p(A), _p(B), _p(C), M, N, K, lda, ldb, ldc, int(a_mn), int(b_mn), epilogue,
                                       0 if block_n == 512 else block_n, fp16, 4, None)
        assert rc == 0, rc

    def norm_fwd(self, x, res_in, w, b, y, res_out, mean, rstd, eps, rms):
        self._count("norm_fwd")
        H = x.size(-1)
        assert self.simt.mlb_norm_fwd(DT[x.dtype], _p(x), _o(res_in), _p(w), _o(b), _p(y), _o(res_out), _o(mean), _p(rstd),
                                      x.numel() // H, H, ctypes.c_float(eps), int(rms), None) == 0

    def norm_bwd(self, dy, x, w, mean, rstd, dres, dx, dw, db, workspace, parts, rms):
        self._count("norm_bwd")
        H = x.size(-1)
        assert self.simt.mlb_norm_bwd(DT[x.dtype], _p(dy), _p(x), _p(w), _o(mean), _p(rstd), _o(dres), _p(dx), _p(dw), _o(db),
                                      _p(workspace), int(parts), x.numel() // H, H, int(rms), None) == 0

    def rope_qkv(self, qkv, freqs, pos, tokens, batch, n_groups, heads_per_group, hn, pos_offset, inverse, token_stride):
        self._count("rope_qkv")
        assert self.simt.mlb_rope_qkv(DT[qkv.dtype], _p(qkv), _p(freqs), _o(pos), tokens, batch, n_groups, heads_per_group, hn,
                                      pos_offset, int(inverse), ctypes.c_longlong(token_stride), None) == 0

    def glu_fwd(self, x, y, kind):
        self._count("glu_fwd")
        F = y.size(-1)
        assert self.simt.mlb_glu_fwd(DT[x.dtype], _p(x), _p(y), ctypes.c_longlong(y.numel() // F), F, kind, None) == 0

    def glu_bwd(self, dy, x, dx, kind):
        self._count("glu_bwd")
        F = dy.size(-1)
        assert self.simt.mlb_glu_bwd(DT[x.dtype], _p(dy), _p(x), _p(dx), ctypes.c_longlong(dy.numel() // F), F, kind, None) == 0

    def gelu(self, x, bias, dy, out, approx, backward):
        self._count("gelu")
        F = x.size(-1)
        assert self.simt.mlb_gelu(DT[x.dtype], _p(x), _o(bias), _o(dy), _p(out), ctypes.c_longlong(x.numel() // F), F,
                                  int(approx), int(backward), None) == 0

    def bias_dropout_add(self, x, bias, residual, out, p, seed, backward):
        self._count("bias_dropout_add")
        F = x.size(-1)
        assert self.simt.mlb_bias_dropout_add(DT[x.dtype], _p(x), _o(bias), _o(residual), _p(out), ctypes.c_longlong(x.numel() // F),
                                              F, ctypes.c_float(p), ctypes.c_ulonglong(seed), int(backward), None) == 0

    def embedding_fwd(self, ids, weight, out, vocab_start, sbh):
        self._count("embedding_fwd")
        assert self.simt.mlb_embedding_fwd(DT[weight.dtype], _p(ids), _p(weight), _p(out), ids.size(0), ids.size(1), weight.size(1),
                                           ctypes.c_longlong(vocab_start), ctypes.c_longlong(weight.size(0)), int(sbh), None) == 0

    def embedding_bwd(self, ids, dout, dweight, vocab_start, sbh):
        self._count("embedding_bwd")
        assert self.simt.mlb_embedding_bwd(DT[dout.dtype], _p(ids), _p(dout), _p(dweight), ids.size(0), ids.size(1), dweight.size(1),
                                           ctypes.c_longlong(vocab_start), ctypes.c_longlong(dweight.size(0)), int(sbh), None) == 0

    def ce_stats(self, logits, target, stats, vocab_start):
        self._count("ce_stats")
        assert self.simt.mlb_ce_stats(DT[logits.dtype], _p(logits), _p(target), _p(stats), logits.size(0), logits.size(1),
                                      int(vocab_start), ctypes.c_longlong(logits.stride(0)), None) == 0

    def ce_bwd(self, logits, out, target, M, logS, g, vocab_start, smoothing, vocab_size):
        self._count("ce_bwd")
        assert self.simt.mlb_ce_bwd(DT[logits.dtype], _p(logits), _p(out), _p(target), _p(M), _p(logS), _p(g), logits.size(0),
                                    logits.size(1), int(vocab_start), ctypes.c_float(smoothing), int(vocab_size),
                                    ctypes.c_longlong(logits.stride(0)), None) == 0

    def adamw_flat(self, p, g, m, v, p16, global_offset, seg_start, seg_wd, seg_lr_mult, lr, beta1, beta2, eps, bc1, bc2,
                   grad_scale, skip, p16_peers):
        self._count("adamw_flat")
        assert not p16_peers
        f = ctypes.c_float
        assert self.simt.mlb_adamw_flat(_p(p), _p(g), _p(m), _p(v), _o(p16), DT[p16.dtype] if p16 is not None else 0,
                                        ctypes.c_longlong(p.numel()), ctypes.c_longlong(global_offset), _p(seg_start), _p(seg_wd),
                                        _o(seg_lr_mult), seg_wd.numel(), f(lr), f(beta1), f(beta2), f(eps), f(bc1), f(bc2),
                                        _o(grad_scale), _o(skip), None, 0, None) == 0

    def sqnorm_flat(self, x, global_offset, seg_start, seg_weight, workspace, out, accumulate):
        self._count("sqnorm_flat")
        assert self.simt.mlb_sqnorm_flat(DT[x.dtype], _p(x), ctypes.c_longlong(x.numel()), ctypes.c_longlong(global_offset),
                                         _o(seg_start), _o(seg_weight), seg_weight.numel() if seg_weight is not None else 0,
                                         _p(workspace), _p(out), int(accumulate), None) == 0

    def clip_coef(self, total_sq, max_norm, norm_out, coef_out, found_inf, extra_scale):
        self._count("clip_coef")
        assert self.simt.mlb_clip_coef(_p(total_sq), ctypes.c_float(max_norm), _p(norm_out), _p(coef_out), _o(found_inf),
                                       ctypes.c_float(extra_scale), None) == 0

    def scale_cast(self, x, y, scale, scale_ptr):
        self._count("scale_cast")
        assert self.simt.mlb_scale_cast(DT[x.dtype], DT[y.dtype], _p(x), _p(y), ctypes.c_longlong(x.numel()), ctypes.c_float(scale),
                                        _o(scale_ptr), None) == 0

    def accumulate(self, x, y):
        self._count("accumulate")
        assert self.simt.mlb_accumulate(DT[x.dtype], _p(x), _p(y), ctypes.c_longlong(x.numel()), None) == 0

    def softmax_fwd(self, x, y, mask, scale, sq, sk, np_, mode):
        self._count("softmax_fwd")
        assert self.simt.mlb_softmax_fwd(DT[x.dtype], _p(x), _p(y), _o(mask), ctypes.c_float(scale), ctypes.c_longlong(x.numel() // sk),
                                         sq, sk, np_, mask.size(0) if mask is not None else 1, mode, None) == 0

    def softmax_bwd(self, dy, y, scale, sk):
        self._count("softmax_bwd")
        assert self.simt.mlb_softmax_bwd(DT[y.dtype], _p(dy), _p(y), ctypes.c_float(scale), ctypes.c_longlong(y.numel() // sk), sk,
                                         None) == 0

    # attention entry points: count them too
    def attn_fwd(self, *a, **k):
        self._count("attn_fwd")
        return super().attn_fwd(*a, **k)

    def attn_bwd(self, *a, **k):
        self._count("attn_bwd")
        return super().attn_bwd(*a, **k)

    def attn_fwd_packed(self, *a, **k):
        self._count("attn_fwd_packed")
        return super().attn_fwd_packed(*a, **k)

    def attn_bwd_packed(self, *a, **k):
        self._count("attn_bwd_packed")
        return super().attn_bwd_packed(*a, **k)
