"""Group algebra + mpu getters (model: reference tests/test_parallel_state.py, but hermetic on Gloo)."""
import pytest
import torch

from megatron_llm_b200.parallel import state as ps
from tests.dist_utils import run_distributed


def test_grid_algebra_matches_reference_layout():
    # world 16, tp 2, pp 4 -> the example in the reference docstring (parallel_state.py:64-87)
    tp, pp, world = 2, 4, 16
    assert ps.tensor_group_ranks(tp, pp, world)[:3] == [[0, 1], [2, 3], [4, 5]]
    assert ps.pipeline_group_ranks(tp, pp, world)[0] == [0, 4, 8, 12]
    assert ps.pipeline_group_ranks(tp, pp, world)[3] == [3, 7, 11, 15]
    assert ps.data_group_ranks(tp, pp, world)[0] == [0, 2]
    assert ps.data_group_ranks(tp, pp, world)[1] == [1, 3]
    assert ps.model_group_ranks(tp, pp, world)[0] == [0, 1, 4, 5, 8, 9, 12, 13]


@pytest.mark.parametrize("tp,pp,world", [(1, 1, 8), (2, 2, 8), (4, 2, 8), (2, 4, 8), (8, 1, 8), (1, 8, 8), (2, 1, 4)])
def test_every_rank_in_exactly_one_group_of_each_family(tp, pp, world):
    for fam in (ps.tensor_group_ranks, ps.pipeline_group_ranks, ps.data_group_ranks, ps.model_group_ranks):
        groups = fam(tp, pp, world)
        flat = sorted(r for g in groups for r in g)
        assert flat == list(range(world))
    for r in range(world):
        p, d, t = ps.grid_coords(r, tp, pp, world)
        assert r == p * (world // pp) + d * tp + t


def _check_getters(rank, world, tp, pp):
    ps.initialize_model_parallel(tp, pp)
    assert ps.model_parallel_is_initialized()
    assert ps.get_tensor_model_parallel_world_size() == tp
    assert ps.get_pipeline_model_parallel_world_size() == pp
    assert ps.get_data_parallel_world_size() == world // (tp * pp)
    assert ps.get_tensor_model_parallel_rank() == rank % tp
    assert ps.get_pipeline_model_parallel_rank() == rank // (world // pp)
    assert ps.get_tensor_model_parallel_src_rank() == (rank // tp) * tp
    assert ps.is_pipeline_first_stage() == (rank // (world // pp) == 0)
    assert ps.is_pipeline_last_stage() == (rank // (world // pp) == pp - 1)
    if pp > 1:
        nxt = ps.get_pipeline_model_parallel_next_rank()
        assert nxt == (rank + world // pp) % world
    # collectives on each group actually work
    t = torch.ones(1)
    torch.distributed.all_reduce(t, group=ps.get_tensor_model_parallel_group())
    assert t.item() == tp
    t = torch.ones(1)
    torch.distributed.all_reduce(t, group=ps.get_data_parallel_group())
    assert t.item() == world // (tp * pp)
    # overrides used by offline tools
    ps.set_tensor_model_parallel_world_size(7)
    assert ps.get_tensor_model_parallel_world_size() == 7
    ps.set_tensor_model_parallel_world_size(None)
    ps.destroy_model_parallel()
    assert not ps.model_parallel_is_initialized()


@pytest.mark.parametrize("tp,pp", [(2, 2), (4, 1), (1, 4)])
def test_getters_world4(tp, pp):
    run_distributed(_check_getters, 4, tp, pp)


def _bad_init(rank, world):
    with pytest.raises(RuntimeError):
        ps.initialize_model_parallel(3, 1)
    with pytest.raises(RuntimeError):
        ps.initialize_model_parallel(1, 2, virtual_pipeline_model_parallel_size=2)


def test_invalid_sizes_raise():
    run_distributed(_bad_init, 2)


def test_fused_tp_graph_replay_bookkeeping():
    """Host side of replaying fused TP kernels from a CUDA graph (parallel/symm.py): the offsets written before a
    replay re-base the captured epochs onto the live counters, the reduce-scatter offset stays even (the receive-slot
    parity is frozen in the captured arguments) and the counters advance by what one replay executes."""
    from megatron_llm_b200.parallel.symm import TPCommunicator

    class FakeMod:
        def __init__(self):
            self.calls = []

        def comm_set_state(self, state, a, b, c):
            self.calls.append((a, b, c))

    comm = object.__new__(TPCommunicator)
    comm.mod, comm.state = FakeMod(), None
    comm.ag_epoch, comm.rs_epoch, comm.rs_arrived_total = 10, 7, 700
    before = comm.counters()
    # a capture pass "issues" 4 all-gather GEMMs and 3 reduce-scatter GEMMs worth 96 arrivals each
    comm.ag_epoch += 4
    comm.rs_epoch += 3
    comm.rs_arrived_total += 288
    advance = comm.end_capture(before)
    assert advance == (4, 3, 288) and comm.counters() == before            # rewound: the capture did not execute
    comm.begin_replay(before, advance)
    assert comm.mod.calls[-1] == (0, 0, 0) and comm.counters() == (14, 10, 988)
    comm.begin_replay(before, advance)                                      # rs delta would be 3 -> one epoch skipped
    assert comm.mod.calls[-1] == (4, 4, 288) and comm.counters() == (18, 14, 1276)
    comm.rs_epoch += 1                                                      # an eager reduce-scatter in between
    comm.rs_arrived_total += 96
    comm.begin_replay(before, advance)
    a, b, c = comm.mod.calls[-1]
    assert b % 2 == 0 and (a, c) == (8, 672) and comm.rs_epoch == before[1] + b + 3


def test_fused_vs_library_selector_follows_the_measured_rates():
    """parallel/fused_tp.py::prefer_fused: at TP=2 the fused kernel on every layer shape of Llama-2-7B; at TP=8 (7 peers:
    350 GB/s in-kernel vs 560 GB/s NCCL-NVLS, measured) only behind the large GEMMs -- the choices the per-pair table
    profiles/fused_tp_n8_r2_graph_mb8.jsonl shows to be the faster ones; the NVLS gather flips the gather pairs back."""
    from megatron_llm_b200.parallel import fused_tp

    class Comm:
        world, nvls_ag, enabled = 8, False, True
    old = fused_tp.communicator()
    try:
        fused_tp.bind(Comm())
        M = 32768                                             # micro-batch 8 x 4096 tokens
        got = {name: fused_tp.prefer_fused(kind, M, n, k) for name, kind, n, k in [
            ("qkv", "ag", 1536, 4096), ("mlp_up", "ag", 2752, 4096), ("attn_dense", "rs", 4096, 512),
            ("dgrad_mlp_down", "ag", 1376, 4096), ("dgrad_attn_dense", "ag", 512, 4096)]}
        assert got == {"qkv": False, "mlp_up": True, "attn_dense": False, "dgrad_mlp_down": False,
                       "dgrad_attn_dense": False}
        Comm.nvls_ag = True
        fused_tp.bind(Comm())                                 # (re-binding clears the decision cache)
        assert all(fused_tp.prefer_fused("ag", M, n, 4096) for n in (1536, 2752, 1376, 512))
        Comm.world, Comm.nvls_ag = 2, False
        fused_tp.bind(Comm())
        assert all(fused_tp.prefer_fused(kind, 8192, n, k) for kind, n, k in [
            ("ag", 6144, 4096), ("ag", 11008, 4096), ("rs", 4096, 2048), ("rs", 4096, 5504), ("ag", 2048, 4096)])
    finally:
        fused_tp.bind(old)
