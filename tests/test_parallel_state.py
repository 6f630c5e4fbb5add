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
