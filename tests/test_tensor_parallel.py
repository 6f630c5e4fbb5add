"""TP mappings, layers, vocab-parallel CE, broadcast_data, RNG tracker on Gloo (model: reference
tests/tensor_parallel/* and mpu/tests/test_layers.py closed-form gradient checks)."""
import pytest
import torch
import torch.distributed as dist

from tests.dist_utils import run_distributed


def _setup(tp):
    from megatron_llm_b200.parallel import state as ps
    ps.initialize_model_parallel(tp, 1)
    from megatron_llm_b200.parallel.random import model_parallel_cuda_manual_seed
    model_parallel_cuda_manual_seed(1234)
    return ps


def _mappings(rank, world):
    ps = _setup(world)
    from megatron_llm_b200.parallel import mappings as m
    x = torch.full((4, 6), float(rank + 1), requires_grad=True)
    # copy: identity fwd, all-reduce bwd
    y = m.copy_to_tensor_model_parallel_region(x)
    y.sum().backward()
    assert torch.equal(y, x) and torch.allclose(x.grad, torch.full_like(x, float(world)))
    # reduce: all-reduce fwd, identity bwd
    x2 = torch.full((4, 6), float(rank + 1), requires_grad=True)
    y = m.reduce_from_tensor_model_parallel_region(x2.clone())
    assert torch.allclose(y, torch.full_like(y, sum(range(1, world + 1))))
    # scatter / gather last dim round trip
    full = torch.arange(24.0).view(4, 6)
    part = m.scatter_to_tensor_model_parallel_region(full)
    assert part.shape == (4, 6 // world)
    assert torch.equal(m.gather_from_tensor_model_parallel_region(part), full)
    # sequence-parallel: scatter-first / gather-first / reduce-scatter
    seq = torch.arange(8.0 * 3).view(8, 3)
    sp = m.scatter_to_sequence_parallel_region(seq)
    assert sp.shape == (8 // world, 3)
    assert torch.equal(m.gather_from_sequence_parallel_region(sp), seq)
    rs = m.reduce_scatter_to_sequence_parallel_region(seq.clone())
    assert torch.allclose(rs, world * seq[rank * (8 // world):(rank + 1) * (8 // world)])
    # gather-first backward = reduce-scatter (tensor-parallel consumer)
    a = torch.ones(2, 3, requires_grad=True)
    g = m.gather_from_sequence_parallel_region(a, True)
    (g * (rank + 1)).sum().backward()
    assert torch.allclose(a.grad, torch.full_like(a, sum(range(1, world + 1))))
    # direct class API (the reference tests call .forward/.backward/.symbolic)
    assert torch.equal(m._CopyToModelParallelRegion.symbolic(None, full), full)
    ps.destroy_model_parallel()


def test_mappings_world2():
    run_distributed(_mappings, 2)


def _layers(rank, world, sequence_parallel):
    ps = _setup(world)
    from megatron_llm_b200.parallel.layers import ColumnParallelLinear, RowParallelLinear, VocabParallelEmbedding
    torch.manual_seed(0)
    s, b, h, out = 8, 2, 16, 24
    X = torch.randn(s, b, h)
    Wc = torch.randn(out, h) * 0.1
    Wr = torch.randn(h, out) * 0.1
    col = ColumnParallelLinear(h, out, bias=False, gather_output=False, use_cpu_initialization=True,
                               sequence_parallel_enabled=sequence_parallel,
                               async_tensor_model_parallel_allreduce=False)
    row = RowParallelLinear(out, h, bias=False, input_is_parallel=True, use_cpu_initialization=True,
                            sequence_parallel_enabled=sequence_parallel)
    with torch.no_grad():
        col.weight.copy_(Wc.chunk(world, 0)[rank])
        row.weight.copy_(Wr.chunk(world, 1)[rank])
    xin = X.chunk(world, 0)[rank].clone() if sequence_parallel else X.clone()
    xin.requires_grad_(True)
    mid, _ = col(xin)
    y, _ = row(torch.tanh(mid))
    # serial reference
    Xr = X.clone().requires_grad_(True)
    Wc_r, Wr_r = Wc.clone().requires_grad_(True), Wr.clone().requires_grad_(True)
    yr = torch.tanh(Xr @ Wc_r.t()) @ Wr_r.t()
    y_full = yr.chunk(world, 0)[rank] if sequence_parallel else yr
    assert torch.allclose(y, y_full, atol=1e-5), (y - y_full).abs().max()
    dy = torch.randn(s, b, h)
    yr.backward(dy)
    y.backward(dy.chunk(world, 0)[rank] if sequence_parallel else dy)
    assert torch.allclose(col.weight.grad, Wc_r.grad.chunk(world, 0)[rank], atol=1e-5)
    assert torch.allclose(row.weight.grad, Wr_r.grad.chunk(world, 1)[rank], atol=1e-5)
    gx = Xr.grad.chunk(world, 0)[rank] if sequence_parallel else Xr.grad
    assert torch.allclose(xin.grad, gx, atol=1e-5)
    # vocab-parallel embedding vs full table
    V = 32
    table = torch.randn(V, h)
    emb = VocabParallelEmbedding(V, h, use_cpu_initialization=True)
    with torch.no_grad():
        emb.weight.copy_(table.chunk(world, 0)[rank])
    ids = torch.randint(0, V, (b, s))
    assert torch.allclose(emb(ids), table[ids], atol=1e-6)
    ps.destroy_model_parallel()


@pytest.mark.parametrize("sp", [False, True])
def test_column_row_vs_serial(sp):
    run_distributed(_layers, 2, sp)


def _gradient_accumulation_fusion(rank, world):
    ps = _setup(world)
    from megatron_llm_b200.parallel.layers import ColumnParallelLinear
    lin = ColumnParallelLinear(8, 12, bias=False, gather_output=False, use_cpu_initialization=True,
                               gradient_accumulation_fusion=True, async_tensor_model_parallel_allreduce=False)
    lin.weight.main_grad = torch.zeros(lin.weight.shape)
    x = torch.randn(4, 2, 8, requires_grad=True)
    for _ in range(2):
        y, _ = lin(x)
        y.sum().backward()
    ref = 2 * torch.ones(8, 12 // world).t() @ x.detach().reshape(-1, 8)
    assert lin.weight.grad is None
    assert torch.allclose(lin.weight.main_grad, ref, atol=1e-5)
    ps.destroy_model_parallel()


def test_wgrad_accumulates_into_main_grad():
    run_distributed(_gradient_accumulation_fusion, 2)


def _cross_entropy(rank, world, smoothing):
    ps = _setup(world)
    from megatron_llm_b200.parallel.cross_entropy import vocab_parallel_cross_entropy, vocab_parallel_max_indices
    torch.manual_seed(5)
    T, V = 12, 16 * world
    logits = torch.randn(3, 4, V) * 3
    target = torch.randint(0, V, (3, 4))
    local = logits.chunk(world, -1)[rank].clone().requires_grad_(True)
    loss = vocab_parallel_cross_entropy(local, target, smoothing)
    ref_in = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in.view(-1, V), target.view(-1), reduction="none",
                                            label_smoothing=0.0).view(3, 4)
    if smoothing > 0:
        lp = torch.log_softmax(ref_in, -1)
        sm = smoothing * V / (V - 1)
        ref = (1 - sm) * ref - sm * lp.mean(-1)
    assert torch.allclose(loss, ref, atol=1e-5), (loss - ref).abs().max()
    w = torch.randn(3, 4)
    (loss * w).sum().backward()
    (ref * w).sum().backward()
    assert torch.allclose(local.grad, ref_in.grad.chunk(world, -1)[rank], atol=1e-5)
    assert torch.equal(vocab_parallel_max_indices(logits.chunk(world, -1)[rank]), logits.argmax(-1))
    ps.destroy_model_parallel()


@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_vocab_parallel_cross_entropy(smoothing):
    run_distributed(_cross_entropy, 2, smoothing)


def _broadcast_data(rank, world):
    ps = _setup(world)
    from megatron_llm_b200.parallel.data import broadcast_data
    keys = ["a", "b"]
    data = {"a": torch.arange(12).view(3, 4), "b": torch.arange(5) + 100} if rank == 0 else None
    out = broadcast_data(keys, data, torch.int64)
    assert torch.equal(out["a"], torch.arange(12).view(3, 4)) and torch.equal(out["b"], torch.arange(5) + 100)
    ps.destroy_model_parallel()


def test_broadcast_data():
    run_distributed(_broadcast_data, 2)


def _rng(rank, world):
    ps = _setup(world)
    from megatron_llm_b200.parallel.random import checkpoint, get_cuda_rng_tracker
    tr = get_cuda_rng_tracker()
    with pytest.raises(Exception):
        tr.add("model-parallel-rng", 99)          # duplicate name
    with tr.fork():
        a = torch.rand(4)
    gathered = [torch.zeros(4) for _ in range(world)]
    dist.all_gather(gathered, a)
    assert not torch.equal(gathered[0], gathered[1])  # TP-forked stream differs across TP ranks
    b = torch.rand(4)                                  # default stream is shared within the TP group
    gathered = [torch.zeros(4) for _ in range(world)]
    dist.all_gather(gathered, b)
    assert torch.equal(gathered[0], gathered[1])

    # recompute replays dropout identically
    lin = torch.nn.Linear(8, 8)

    def fn(x):
        with tr.fork():
            return torch.nn.functional.dropout(lin(x), 0.5, training=True)

    x1 = torch.randn(4, 8, requires_grad=True)
    states = tr.get_states()
    y = checkpoint(fn, False, x1)
    y.sum().backward()
    tr.set_states(states)
    x2 = x1.detach().clone().requires_grad_(True)
    lin.zero_grad()
    y2 = fn(x2)
    y2.sum().backward()
    assert torch.allclose(y, y2) and torch.allclose(x1.grad, x2.grad)
    ps.destroy_model_parallel()


def test_rng_tracker_and_checkpoint():
    run_distributed(_rng, 2)
