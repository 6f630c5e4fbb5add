"""Forward/backward schedules: no pipelining, 1F1B, interleaved 1F1B.

Parity target: megatron/schedules.py (get_forward_backward_func :18-34, deallocate_output_tensor :36-53,
custom_backward :55-88, forward_step :91-139, backward_step :142-202, no-pipelining :213-250, interleaved
:253-502, get_tensor_shapes :505-535, 1F1B :606-722).

Differences: P2P receives are stream-ordered (no device synchronise per transfer, see parallel/p2p.py); the
data-parallel wrapper is told when the last micro-batch's backward starts (``enable_grad_sync``) so bucketed
gradient reduction overlaps it.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch
from torch.autograd.variable import Variable

from ..models.enums import ModelType
from . import p2p
from . import state as ps


def _glob():
    from ..global_vars import get_args, get_num_microbatches, get_timers
    return get_args, get_num_microbatches, get_timers


def _unwrap(model):
    from ..utils import unwrap_model
    return unwrap_model(model)


def get_forward_backward_func():
    get_args, get_num_microbatches, _ = _glob()
    args = get_args()
    if ps.get_pipeline_model_parallel_world_size() > 1:
        if args.virtual_pipeline_model_parallel_size is not None:
            assert get_num_microbatches() % args.pipeline_model_parallel_size == 0, \
                "number of microbatches (%d) is not divisible by pipeline-model-parallel-size (%d) when using " \
                "interleaved schedule" % (get_num_microbatches(), args.pipeline_model_parallel_size)
            return forward_backward_pipelining_with_interleaving
        return forward_backward_pipelining_without_interleaving
    return forward_backward_no_pipelining


def deallocate_output_tensor(out):
    """After an activation has been sent downstream only its ``.grad_fn`` is needed: shrink ``.data`` to one
    element so the memory is released while the autograd graph stays alive."""
    if out is None:
        return
    assert isinstance(out, torch.Tensor), "expected Tensor, found %s." % type(out).__name__
    assert out._base is None, "counter-productive to free a view of another tensor."
    out.data = torch.empty((1,), device=out.device, dtype=out.dtype)


def custom_backward(output, grad_output):
    """``torch.autograd.backward`` minus the output/grad shape check (the output was deallocated above)."""
    assert output.numel() == 1, "output should be pseudo-'freed' in schedule, to optimize memory"
    assert isinstance(output, torch.Tensor), "output == '%s'." % type(output).__name__
    assert isinstance(grad_output, (torch.Tensor, type(None))), "grad_output == '%s'." % type(grad_output).__name__
    if grad_output is None:
        assert output.numel() == 1, "implicit grad requires scalar output."
        grad_output = torch.ones_like(output, memory_format=torch.preserve_format)
    Variable._execution_engine.run_backward(tensors=(output,), grad_tensors=(grad_output,), keep_graph=False,
                                            create_graph=False, inputs=tuple(), allow_unreachable=True,
                                            accumulate_grad=True)


def forward_step(forward_step_func, data_iterator, model, input_tensor, forward_data_store, timers,
                 collect_non_loss_data=False):
    """Returns the stage output (the scaled loss on the last stage); loss dicts go to ``forward_data_store``."""
    get_args, get_num_microbatches, _ = _glob()
    args = get_args()
    if timers is not None:
        timers("forward-compute", log_level=2).start()
    unwrapped = _unwrap(model)
    unwrap_output_tensor = False
    if not isinstance(input_tensor, list):
        input_tensor = [input_tensor]
        unwrap_output_tensor = True
    unwrapped.set_input_tensor(input_tensor)
    output_tensor, loss_func = forward_step_func(data_iterator, model)
    if ps.is_pipeline_last_stage():
        if not collect_non_loss_data:
            output_tensor = loss_func(output_tensor)
            loss, loss_reduced = output_tensor
            output_tensor = loss / get_num_microbatches()
            forward_data_store.append(loss_reduced)
        else:
            forward_data_store.append(loss_func(output_tensor, non_loss_data=True))
    if timers is not None:
        timers("forward-compute").stop()
    if ps.is_pipeline_stage_after_split() and args.model_type == ModelType.encoder_and_decoder:
        return [output_tensor, input_tensor[-1]]
    if unwrap_output_tensor:
        return output_tensor
    return [output_tensor]


def backward_step(optimizer, input_tensor, output_tensor, output_tensor_grad, timers):
    """Backward through this stage; returns d(loss)/d(input_tensor) (None on the first stage)."""
    get_args, _, _ = _glob()
    args = get_args()
    if timers is not None:
        timers("backward-compute", log_level=2).start()
    unwrap_input_tensor_grad = False
    if not isinstance(input_tensor, list):
        input_tensor = [input_tensor]
        unwrap_input_tensor_grad = True
    for x in input_tensor:
        if x is not None:
            x.retain_grad()
    if not isinstance(output_tensor, list):
        output_tensor = [output_tensor]
    if not isinstance(output_tensor_grad, list):
        output_tensor_grad = [output_tensor_grad]
    if output_tensor_grad[0] is None and optimizer is not None:
        output_tensor = optimizer.scale_loss(output_tensor[0])
        output_tensor = [output_tensor]
    if output_tensor_grad[0] is None:
        torch.autograd.backward(output_tensor[0], grad_tensors=None)
    else:
        custom_backward(output_tensor[0], output_tensor_grad[0])
    input_tensor_grad = [None]
    if input_tensor is not None:
        input_tensor_grad = [None if x is None else x.grad for x in input_tensor]
    # skip connection of the encoder output in encoder-decoder pipelines
    if ps.get_pipeline_model_parallel_world_size() > 1 and ps.is_pipeline_stage_after_split() and \
            args.model_type == ModelType.encoder_and_decoder:
        if output_tensor_grad[1] is not None:
            input_tensor_grad[-1].add_(output_tensor_grad[1])
    if unwrap_input_tensor_grad:
        input_tensor_grad = input_tensor_grad[0]
    if timers is not None:
        timers("backward-compute").stop()
    return input_tensor_grad


@contextmanager
def dummy_handler():
    yield


def _enable_grad_sync(model, flag=True):
    models = model if isinstance(model, list) else [model]
    for m in models:
        fn = getattr(m, "enable_grad_sync", None)
        if fn is not None:
            fn(flag)


class GraphedMicrobatch:
    """One micro-batch (forward + loss + backward) captured in a CUDA graph and replayed.

    With tensor parallelism the per-GPU kernels shrink by 1/TP while the host still issues the same ~2.7k launches per
    micro-batch, so the step becomes launch-bound (``host_enqueue_ms_per_step`` in bench.py).  Replaying a graph
    removes the host from the loop.  Requirements (checked by :func:`_graph_runner`): no pipeline parallelism, DP = 1
    (gradient-bucket hooks are Python), no dropout / activation recompute (host-side RNG bookkeeping), static shapes,
    and a forward_step_func that exposes its two halves as ``.get_batch(data_iterator)`` and ``.run(batch, model)``.
    Gradients accumulate into the static fp32 ``main_grad`` buffers exactly as in eager mode (the accumulation hooks'
    kernels are part of the graph)."""

    WARMUP = 2

    def __init__(self, parts, model, optimizer):
        self.parts, self.model, self.optimizer = parts, model, optimizer
        self.graph = None
        self.calls = 0
        self.static_batch = None
        self.static_reduced = None
        self.launches = 0
        self.comm = None

    def _eager(self, batch, n_microbatches):
        output, loss_func = self.parts.run(batch, self.model)
        loss, reduced = loss_func(output)
        scaled = loss / n_microbatches
        if self.optimizer is not None:
            scaled = self.optimizer.scale_loss(scaled)
        torch.autograd.backward(scaled)
        return reduced

    def run(self, batch, n_microbatches):
        from ..ops import _ext
        self.calls += 1
        if self.graph is None and self.calls <= self.WARMUP:
            return self._eager(batch, n_microbatches)
        if self.graph is None:
            torch.cuda.synchronize()
            self.static_batch = [b.clone() if torch.is_tensor(b) else b for b in batch]
            self.graph = torch.cuda.CUDAGraph()
            before = _ext.LAUNCHES
            from . import fused_tp
            self.comm = fused_tp.communicator()
            if self.comm is not None:
                self.comm_before = self.comm.counters()
            with torch.cuda.graph(self.graph, stream=torch.cuda.current_stream(), capture_error_mode="thread_local"):
                self.static_reduced = self._eager(self.static_batch, n_microbatches)
            if self.comm is not None:
                self.comm_advance = self.comm.end_capture(self.comm_before)
            self.launches = _ext.LAUNCHES - before
            torch.cuda.synchronize()
            # the capture pass itself does not execute: fall through and replay for this micro-batch
        for dst, src in zip(self.static_batch, batch):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
        if self.comm is not None:   # fused TP kernels: re-base the captured epochs onto the live sequence
            self.comm.begin_replay(self.comm_before, self.comm_advance)
        self.graph.replay()
        _ext.count(self.launches)
        return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.static_reduced.items()}


_GRAPH_RUNNERS = {}


def _graph_runner(forward_step_func, model, optimizer, forward_only):
    get_args, _, _ = _glob()
    args = get_args()
    if forward_only or not getattr(args, "cuda_graph_microbatch", False) or not torch.cuda.is_available():
        return None
    if not (hasattr(forward_step_func, "get_batch") and hasattr(forward_step_func, "run")):
        return None
    ok = (ps.get_data_parallel_world_size() == 1 and args.hidden_dropout == 0.0 and args.attention_dropout == 0.0
          and args.recompute_granularity is None and not args.variable_seq_lengths and not args.fp16)
    if not ok:
        if not getattr(_graph_runner, "warned", False):
            print("WARNING: --cuda_graph_microbatch ignored (needs DP=1, no dropout, no recompute, bf16/fp32, fixed "
                  "sequence length)", flush=True)
            _graph_runner.warned = True
        return None
    key = id(model)
    if key not in _GRAPH_RUNNERS:
        _GRAPH_RUNNERS[key] = GraphedMicrobatch(forward_step_func, model, optimizer)
    return _GRAPH_RUNNERS[key]


def forward_backward_no_pipelining(forward_step_func, data_iterator, model, optimizer, timers, forward_only,
                                   collect_non_loss_data=False):
    """All micro-batches forward+backward on one stage; DP grad reduction overlaps the last backward."""
    _, get_num_microbatches, _ = _glob()
    assert len(model) == 1
    model = model[0]
    runner = _graph_runner(forward_step_func, model, optimizer, forward_only) if not collect_non_loss_data else None
    if runner is not None:
        n = get_num_microbatches()
        _enable_grad_sync(model, True)
        return [runner.run(forward_step_func.get_batch(data_iterator), n) for _ in range(n)]
    from torch.nn.parallel.distributed import DistributedDataParallel as torchDDP
    context_handler = model.no_sync if isinstance(model, torchDDP) else dummy_handler
    forward_data_store = []
    input_tensor, output_tensor_grad = None, None
    n = get_num_microbatches()
    with context_handler():
        for _ in range(n - 1):
            output_tensor = forward_step(forward_step_func, data_iterator, model, input_tensor, forward_data_store,
                                         timers, collect_non_loss_data)
            if not forward_only:
                backward_step(optimizer, input_tensor, output_tensor, output_tensor_grad, timers)
    output_tensor = forward_step(forward_step_func, data_iterator, model, input_tensor, forward_data_store, timers,
                                 collect_non_loss_data)
    if not forward_only:
        _enable_grad_sync(model, True)
        backward_step(optimizer, input_tensor, output_tensor, output_tensor_grad, timers)
    return forward_data_store


def get_tensor_shapes(rank, model_type):
    """Shapes of the tensors exchanged between stages: [s (/tp under SP), b, h]; encoder-decoder stages after the
    split exchange (decoder, encoder) pairs."""
    get_args, _, _ = _glob()
    args = get_args()
    tp = ps.get_tensor_model_parallel_world_size()
    seq = args.seq_length // tp if args.sequence_parallel else args.seq_length
    shapes = []
    if model_type == ModelType.encoder_and_decoder:
        dec = args.decoder_seq_length // tp if args.sequence_parallel else args.decoder_seq_length
        if ps.is_pipeline_stage_before_split(rank):
            shapes.append((seq, args.micro_batch_size, args.hidden_size))
        else:
            shapes.append((dec, args.micro_batch_size, args.hidden_size))
            shapes.append((seq, args.micro_batch_size, args.hidden_size))
    else:
        shapes.append((seq, args.micro_batch_size, args.hidden_size))
    return shapes


def recv_forward(tensor_shapes, timers):
    return [None if s is None else p2p.recv_forward(s, timers=timers) for s in tensor_shapes]


def recv_backward(tensor_shapes, timers):
    return [None if s is None else p2p.recv_backward(s, timers=timers) for s in tensor_shapes]


def send_forward(output_tensors, tensor_shapes, timers):
    if not isinstance(output_tensors, list):
        output_tensors = [output_tensors]
    for out, s in zip(output_tensors, tensor_shapes):
        if s is not None:
            p2p.send_forward(out, s, timers=timers)


def send_backward(input_tensor_grads, tensor_shapes, timers):
    if not isinstance(input_tensor_grads, list):
        input_tensor_grads = [input_tensor_grads]
    for g, s in zip(input_tensor_grads, tensor_shapes):
        if s is not None:
            p2p.send_backward(g, s, timers=timers)


def send_forward_recv_backward(output_tensors, tensor_shapes, timers):
    if not isinstance(output_tensors, list):
        output_tensors = [output_tensors]
    return [None if s is None else p2p.send_forward_recv_backward(out, s, timers=timers)
            for out, s in zip(output_tensors, tensor_shapes)]


def send_backward_recv_forward(input_tensor_grads, tensor_shapes, timers):
    if not isinstance(input_tensor_grads, list):
        input_tensor_grads = [input_tensor_grads]
    return [None if s is None else p2p.send_backward_recv_forward(g, s, timers=timers)
            for g, s in zip(input_tensor_grads, tensor_shapes)]


def forward_backward_pipelining_without_interleaving(forward_step_func, data_iterator, model, optimizer, timers,
                                                     forward_only, collect_non_loss_data=False):
    """Non-interleaved 1F1B: warm-up forwards, steady one-forward-one-backward, cool-down backwards."""
    get_args, get_num_microbatches, _ = _glob()
    args = get_args()
    assert len(model) == 1
    model = model[0]
    n_mb = get_num_microbatches()
    pp_world, pp_rank = ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_rank()
    num_warmup = min(pp_world - pp_rank - 1, n_mb)
    num_remaining = n_mb - num_warmup
    model_type = _unwrap(model).model_type if hasattr(_unwrap(model), "model_type") else args.model_type
    recv_shapes = get_tensor_shapes(pp_rank - 1, model_type)
    send_shapes = get_tensor_shapes(pp_rank, model_type)
    input_tensors, output_tensors = ([], []) if not forward_only else (None, None)
    forward_data_store = []
    total_backwards = n_mb
    done_backwards = 0

    def run_backward(inp, out, out_grad):
        nonlocal done_backwards
        done_backwards += 1
        if done_backwards == total_backwards:
            _enable_grad_sync(model, True)
        return backward_step(optimizer, inp, out, out_grad, timers)

    for _ in range(num_warmup):
        input_tensor = recv_forward(recv_shapes, timers)
        output_tensor = forward_step(forward_step_func, data_iterator, model, input_tensor, forward_data_store,
                                     timers, collect_non_loss_data)
        send_forward(output_tensor, send_shapes, timers)
        if not forward_only:
            input_tensors.append(input_tensor)
            output_tensors.append(output_tensor)
            deallocate_output_tensor(output_tensor[0])
    if num_remaining > 0:
        input_tensor = recv_forward(recv_shapes, timers)
    for i in range(num_remaining):
        last_iteration = i == (num_remaining - 1)
        output_tensor = forward_step(forward_step_func, data_iterator, model, input_tensor, forward_data_store,
                                     timers, collect_non_loss_data)
        if forward_only:
            send_forward(output_tensor, send_shapes, timers)
            if not last_iteration:
                input_tensor = recv_forward(recv_shapes, timers)
        else:
            output_tensor_grad = send_forward_recv_backward(output_tensor, send_shapes, timers)
            input_tensors.append(input_tensor)
            output_tensors.append(output_tensor)
            deallocate_output_tensor(output_tensor[0])
            input_tensor, output_tensor = input_tensors.pop(0), output_tensors.pop(0)
            input_tensor_grad = run_backward(input_tensor, output_tensor, output_tensor_grad)
            if last_iteration:
                input_tensor = None
                send_backward(input_tensor_grad, recv_shapes, timers)
            else:
                input_tensor = send_backward_recv_forward(input_tensor_grad, recv_shapes, timers)
    if not forward_only:
        for _ in range(num_warmup):
            input_tensor, output_tensor = input_tensors.pop(0), output_tensors.pop(0)
            output_tensor_grad = recv_backward(send_shapes, timers)
            input_tensor_grad = run_backward(input_tensor, output_tensor, output_tensor_grad)
            send_backward(input_tensor_grad, recv_shapes, timers)
    p2p.drain_pending_sends()
    return forward_data_store


def forward_backward_pipelining_with_interleaving(forward_step_func, data_iterator, model, optimizer, timers,
                                                  forward_only, collect_non_loss_data=False):
    """Interleaved 1F1B: every rank owns ``len(model)`` virtual stages (model chunks); micro-batches are
    processed in groups of ``pp`` per chunk so the pipeline bubble shrinks by the number of chunks."""
    get_args, get_num_microbatches, _ = _glob()
    args = get_args()
    num_chunks = len(model)
    input_tensors = [[] for _ in range(num_chunks)]
    output_tensors = [[] for _ in range(num_chunks)]
    forward_data_store = []
    output_tensor_grads = [[] for _ in range(num_chunks)] if not forward_only else None
    pp_size, pp_rank = ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_rank()
    seq = args.seq_length // ps.get_tensor_model_parallel_world_size() if args.sequence_parallel else args.seq_length
    tensor_shape = (seq, args.micro_batch_size, args.hidden_size)
    n_mb = get_num_microbatches() * num_chunks
    all_warmup = False
    if forward_only:
        num_warmup = n_mb
    elif get_num_microbatches() == pp_size:
        num_warmup, all_warmup = n_mb, True
    else:
        num_warmup = min((pp_size - pp_rank - 1) * 2 + (num_chunks - 1) * pp_size, n_mb)
    num_remaining = n_mb - num_warmup
    done_backwards = 0

    def chunk_id(k, forward):
        in_group = k % (pp_size * num_chunks)
        c = in_group // pp_size
        return c if forward else (num_chunks - c - 1)

    def forward_helper(k):
        c = chunk_id(k, True)
        ps.set_virtual_pipeline_model_parallel_rank(c)
        if ps.is_pipeline_first_stage() and len(input_tensors[c]) == len(output_tensors[c]):
            input_tensors[c].append(None)
        out = forward_step(forward_step_func, data_iterator[c], model[c], input_tensors[c][-1],
                           forward_data_store, timers, collect_non_loss_data)
        output_tensors[c].append(out)
        if forward_only:
            input_tensors[c].pop()
            output_tensors[c].pop()
        return out

    def backward_helper(k):
        nonlocal done_backwards
        c = chunk_id(k, False)
        ps.set_virtual_pipeline_model_parallel_rank(c)
        if ps.is_pipeline_last_stage() and len(output_tensor_grads[c]) == 0:
            output_tensor_grads[c].append(None)
        done_backwards += 1
        if done_backwards == n_mb:
            _enable_grad_sync(model, True)
        return backward_step(optimizer, input_tensors[c].pop(0), output_tensors[c].pop(0),
                             output_tensor_grads[c].pop(0), timers)

    ps.set_virtual_pipeline_model_parallel_rank(0)
    input_tensors[0].append(p2p.recv_forward(tensor_shape, timers=timers))
    for k in range(num_warmup):
        output_tensor = forward_helper(k)
        next_c = chunk_id(k + 1, True)
        recv_prev = True
        if ps.is_pipeline_first_stage(ignore_virtual=True) and next_c == 0:
            recv_prev = False
        if k == n_mb - 1:
            recv_prev = False
        if ps.is_pipeline_last_stage():
            output_tensor = None
        if k == num_warmup - 1 and not forward_only and not all_warmup:
            input_tensor_grad = None
            recv_next = not ps.is_pipeline_last_stage(ignore_virtual=True)
            input_tensor, output_tensor_grad = p2p.send_forward_backward_recv_forward_backward(
                output_tensor, input_tensor_grad, recv_prev=recv_prev, recv_next=recv_next,
                tensor_shape=tensor_shape, timers=timers)
            output_tensor_grads[num_chunks - 1].append(output_tensor_grad)
        else:
            input_tensor = p2p.send_forward_recv_forward(output_tensor, recv_prev=recv_prev,
                                                         tensor_shape=tensor_shape, timers=timers)
        input_tensors[next_c].append(input_tensor)
        deallocate_output_tensor(output_tensor)

    for k in range(num_remaining):
        fk = k + num_warmup
        output_tensor = forward_helper(fk)
        input_tensor_grad = backward_helper(k)
        fc = chunk_id(fk, True)
        ps.set_virtual_pipeline_model_parallel_rank(fc)
        if ps.is_pipeline_last_stage():
            output_tensor = None
        bc = chunk_id(k, False)
        ps.set_virtual_pipeline_model_parallel_rank(bc)
        if ps.is_pipeline_first_stage():
            input_tensor_grad = None
        recv_prev = True
        if ps.is_pipeline_first_stage(ignore_virtual=True):
            next_fc = chunk_id(fk - (pp_size - 1), True)
            if next_fc == num_chunks - 1:
                recv_prev = False
            next_fc += 1
        else:
            next_fc = chunk_id(fk + 1, True)
        recv_next = True
        if ps.is_pipeline_last_stage(ignore_virtual=True):
            next_bc = chunk_id(k - (pp_size - 1), False)
            if next_bc == 0:
                recv_next = False
            next_bc -= 1
        else:
            next_bc = chunk_id(k + 1, False)
        if k == num_remaining - 1:
            recv_prev = False
        input_tensor, output_tensor_grad = p2p.send_forward_backward_recv_forward_backward(
            output_tensor, input_tensor_grad, recv_prev=recv_prev, recv_next=recv_next, tensor_shape=tensor_shape,
            timers=timers)
        deallocate_output_tensor(output_tensor)
        if recv_prev:
            input_tensors[next_fc].append(input_tensor)
        if recv_next:
            output_tensor_grads[next_bc].append(output_tensor_grad)

    if not forward_only:
        if all_warmup:
            output_tensor_grads[num_chunks - 1].append(p2p.recv_backward(tensor_shape, timers=timers))
        for k in range(num_remaining, n_mb):
            input_tensor_grad = backward_helper(k)
            next_bc = chunk_id(k + 1, False)
            recv_next = True
            if ps.is_pipeline_last_stage(ignore_virtual=True) and next_bc == (num_chunks - 1):
                recv_next = False
            if k == n_mb - 1:
                recv_next = False
            output_tensor_grads[next_bc].append(p2p.send_backward_recv_backward(
                input_tensor_grad, recv_next=recv_next, tensor_shape=tensor_shape, timers=timers))
    p2p.drain_pending_sends()
    return forward_data_store
