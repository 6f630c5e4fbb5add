"""GPU check of the packed-QKV attention path (fwd + bwd) against the fp32 reference on the split views."""
import faulthandler, json, os, sys
faulthandler.enable()
import torch
sys.path.insert(0, ".")
from megatron_llm_b200.ops import attention_sm100 as A
from megatron_llm_b200.ops.attention import attention_reference

def run(b, s, nkv, g, window, hn=128):
    torch.manual_seed(0)
    mixed = (torch.randn(s, b, nkv * (g + 2) * hn, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    assert A.packed_supported(mixed, nkv, g, hn, 0.0)
    out = A.packed_attention(mixed, nkv, g, window, None, hn)
    mr = mixed.detach().float().requires_grad_(True)
    qkv = mr.view(s, b, nkv, g + 2, hn)
    q = qkv[:, :, :, :g].reshape(s, b, nkv * g, hn).transpose(0, 1)
    k = qkv[:, :, :, g].transpose(0, 1)
    v = qkv[:, :, :, g + 1].transpose(0, 1)
    ref = attention_reference(q, k, v, causal=True, window=window).transpose(0, 1).reshape(s, b, -1)
    do = torch.randn_like(out)
    out.backward(do)
    ref.backward(do.float())
    res = {"packed": True, "b": b, "s": s, "nkv": nkv, "g": g, "hn": hn, "window": window,
           "fwd_err": (out.float() - ref).abs().max().item(), "fwd_scale": ref.abs().max().item(),
           "dmixed_err": (mixed.grad.float() - mr.grad).abs().max().item(), "dmixed_scale": mr.grad.abs().max().item()}
    print(json.dumps(res), flush=True)

a = sys.argv[1:]
hn = int(a[5]) if len(a) > 5 else 128
if hn == 64:
    os.environ["MLB200_ATTN_HD64"] = "1"
run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), None if a[4] == "none" else int(a[4]), hn)
