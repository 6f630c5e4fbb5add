"""CPU oracles of the hot ops: checks the reference implementations (used for CPU/Gloo runs and as the GPU
kernels' oracle) against independent formulas, incl. their hand-written backward passes."""
import pytest
import torch

from megatron_llm_b200 import ops


def test_rmsnorm_backward_matches_autograd():
    torch.manual_seed(0)
    x = torch.randn(5, 32, requires_grad=True)
    w = torch.randn(32, requires_grad=True)
    r = torch.randn(5, 32, requires_grad=True)
    y, h = ops.rmsnorm(x, w, 1e-5, residual=r)
    x2, w2, r2 = (t.detach().clone().requires_grad_(True) for t in (x, w, r))
    h2 = x2 + r2
    y2 = h2 * torch.rsqrt(h2.pow(2).mean(-1, keepdim=True) + 1e-5) * w2
    dy, dh = torch.randn_like(y), torch.randn_like(h)
    torch.autograd.backward([y, h], [dy, dh])
    torch.autograd.backward([y2, h2], [dy, dh])
    for a, b in ((x.grad, x2.grad), (w.grad, w2.grad), (r.grad, r2.grad)):
        assert torch.allclose(a, b, atol=1e-5)


def test_layernorm_backward_matches_autograd():
    torch.manual_seed(1)
    x = torch.randn(6, 16, requires_grad=True)
    w, b = torch.randn(16, requires_grad=True), torch.randn(16, requires_grad=True)
    y = ops.layernorm(x, w, b, 1e-5)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y2 = torch.nn.functional.layer_norm(x2, (16,), w2, b2, 1e-5)
    dy = torch.randn_like(y)
    y.backward(dy)
    y2.backward(dy)
    assert torch.allclose(y, y2, atol=1e-5)
    for a, c in ((x.grad, x2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        assert torch.allclose(a, c, atol=1e-5)


def test_rope_reference_equals_complex_formulation():
    from megatron_llm_b200.models.positional_embeddings import apply_rotary_emb, precompute_freqs_cis
    torch.manual_seed(2)
    s, b, nkv, g, hn = 16, 2, 2, 2, 8
    mixed = torch.randn(s, b, nkv * (g + 2) * hn)
    table = ops.rope_table(hn, 32, 10000.0, 1.0)
    pos = torch.randint(0, 32, (b, s))
    ref = mixed.clone().view(s, b, nkv, g + 2, hn)
    out = ops.rope_qkv_(mixed.clone(), nkv, g + 2, hn, table, pos, 0).view(s, b, nkv, g + 2, hn)
    fc = precompute_freqs_cis(hn, 32)
    q, k = apply_rotary_emb(ref[:, :, :, :g].reshape(s, b, nkv * g, hn), ref[:, :, :, g], fc, position_ids=pos)
    assert torch.allclose(out[:, :, :, :g].reshape(s, b, nkv * g, hn), q, atol=1e-5)
    assert torch.allclose(out[:, :, :, g], k, atol=1e-5)
    assert torch.equal(out[:, :, :, g + 1], ref[:, :, :, g + 1])


@pytest.mark.parametrize("kind", ["swiglu", "geglu", "reglu", "liglu"])
def test_glu_matches_reference_definition(kind):
    """Same oracle as the reference's tests/test_activations.py: x1 * act(x2) with chunk(2, -1)."""
    x = torch.randn(3, 10)
    x1, x2 = x.chunk(2, -1)
    act = {"swiglu": torch.nn.functional.silu, "geglu": torch.nn.functional.gelu, "reglu": torch.relu,
           "liglu": lambda z: z}[kind]
    assert torch.allclose(ops.glu(x, kind), x1 * act(x2), atol=1e-6)


def test_attention_reference_gqa_and_window():
    from megatron_llm_b200.ops.attention import attention_reference
    torch.manual_seed(3)
    b, s, nq, nkv, hn = 1, 12, 4, 2, 8
    q, k, v = torch.randn(b, s, nq, hn), torch.randn(b, s, nkv, hn), torch.randn(b, s, nkv, hn)
    out = attention_reference(q, k, v, causal=True)
    ke, ve = k.repeat_interleave(2, dim=2), v.repeat_interleave(2, dim=2)
    ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), ke.transpose(1, 2), ve.transpose(1, 2),
                                                           is_causal=True).transpose(1, 2)
    assert torch.allclose(out, ref, atol=1e-5)
    outw = attention_reference(q, k, v, causal=True, window=3)
    assert not torch.allclose(outw, out)
    # first rows (fewer than window keys) are unaffected by the window
    assert torch.allclose(outw[:, :3], out[:, :3], atol=1e-6)


def test_softmax_reference_modes():
    x = torch.randn(2, 2, 5, 5)
    y = ops.scaled_upper_triang_masked_softmax(x, 1.0)
    assert torch.allclose(y.sum(-1), torch.ones(2, 2, 5), atol=1e-5)
    assert (y.triu(1) == 0).all()
    mask = torch.zeros(2, 1, 5, 5, dtype=torch.bool)
    mask[:, :, 0] = True                      # fully masked row -> zeros
    ym = ops.scaled_masked_softmax(x, mask, 1.0)
    assert (ym[:, :, 0] == 0).all()
