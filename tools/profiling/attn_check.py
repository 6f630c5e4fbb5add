"""GPU check of the tcgen05 attention kernels against the fp32 reference (+ timing vs the FA-2 library).

  python tools/profiling/attn_check.py B S NQ NKV WINDOW|none c|t [HN=128]      (c = correctness, t = timing)
"""
import faulthandler, json, os, sys, time
faulthandler.enable()
_t0 = time.time()
import torch
def log(m):
    print(f"[{time.time()-_t0:7.1f}s] {m}", file=sys.stderr, flush=True)
log("torch imported")
sys.path.insert(0, ".")
from megatron_llm_b200.ops import attention_sm100 as A
from megatron_llm_b200.ops.attention import attention_reference

def run(b, s, nq, nkv, window, timing, hn=128):
    torch.manual_seed(0)
    dev = "cuda"
    q = (torch.randn(b, s, nq, hn, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    k = (torch.randn(b, s, nkv, hn, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    v = (torch.randn(b, s, nkv, hn, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    assert A.supported(q, k, v, True, window, 0.0), "kernel not available / shape unsupported"
    log("inputs ready")
    out = A.attention(q, k, v, True, window, None)
    torch.cuda.synchronize(); log("fwd done")
    res = {"b": b, "s": s, "nq": nq, "nkv": nkv, "hn": hn, "window": window}
    if not timing:
        qr, kr, vr = (t.detach().float().requires_grad_(True) for t in (q, k, v))
        ref = attention_reference(qr, kr, vr, causal=True, window=window)
        log("reference fwd done")
        res["fwd_err"] = (out.float() - ref).abs().max().item()
        res["fwd_ref_scale"] = ref.abs().max().item()
        do = torch.randn_like(out)
        out.backward(do)
        torch.cuda.synchronize(); log("bwd done")
        ref.backward(do.float())
        for name, a, bb in (("dq", q.grad, qr.grad), ("dk", k.grad, kr.grad), ("dv", v.grad, vr.grad)):
            res[name + "_err"] = (a.float() - bb).abs().max().item()
            res[name + "_scale"] = bb.abs().max().item()
    else:
        from flash_attn import flash_attn_func
        do = torch.randn_like(out)
        def t(fn, n=10):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record(); fn(); e_.record(); torch.cuda.synchronize(); ts.append(s_.elapsed_time(e_))
            return sorted(ts)[len(ts) // 2]
        fl = 2.0 * b * nq * s * s * hn * 2 * 0.5   # causal fwd flops
        res["fwd_ms"] = t(lambda: A.attention(q, k, v, True, window, None))
        res["fa2_fwd_ms"] = t(lambda: flash_attn_func(q, k, v, causal=True))
        def fb_mine():
            q.grad = k.grad = v.grad = None
            A.attention(q, k, v, True, window, None).backward(do)
        def fb_fa():
            q.grad = k.grad = v.grad = None
            flash_attn_func(q, k, v, causal=True).backward(do)
        res["fwdbwd_ms"] = t(fb_mine); res["fa2_fwdbwd_ms"] = t(fb_fa)
        res["fwd_tflops"] = fl / res["fwd_ms"] / 1e9; res["fa2_fwd_tflops"] = fl / res["fa2_fwd_ms"] / 1e9
        res["fwdbwd_tflops"] = 3.5 * fl / res["fwdbwd_ms"] / 1e9; res["fa2_fwdbwd_tflops"] = 3.5 * fl / res["fa2_fwdbwd_ms"] / 1e9
    print(json.dumps(res), flush=True)

args = sys.argv[1:]
hn = int(args[6]) if len(args) > 6 else 128
if hn == 64:
    os.environ["MLB200_ATTN_HD64"] = "1"      # (default since round 2)
run(int(args[0]), int(args[1]), int(args[2]), int(args[3]), None if args[4] == "none" else int(args[4]), args[5] == "t", hn)
