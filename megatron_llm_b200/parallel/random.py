"""Model-parallel RNG tracker and activation recompute.

Parity target: megatron/core/tensor_parallel/random.py (:64-132 tracker, :144-172 seeding,
:175-252 CheckpointFunction incl. ``distribute_saved_activations``).

Device-agnostic: on a B200 the tracked state is the CUDA generator state; on CPU/Gloo it is
the CPU generator state, so dropout replay semantics are testable without a GPU.
"""
from __future__ import annotations

import contextlib

import torch
from torch.utils.checkpoint import detach_variable

from ..utils.core_utils import safely_set_viewless_tensor_data
from ..utils.device import use_cuda
from . import state as ps
from .tp_utils import gather_split_1d_tensor, split_tensor_into_1d_equal_chunks

_MODEL_PARALLEL_RNG_TRACKER_NAME = "model-parallel-rng"


def _get_device_rng_state():
    return torch.cuda.get_rng_state() if use_cuda() else torch.get_rng_state()


def _set_device_rng_state(state):
    if use_cuda():
        # default_generators[idx].set_state avoids the device-sync + clone of the public API
        idx = torch.cuda.current_device()
        torch.cuda.default_generators[idx].set_state(state)
    else:
        torch.set_rng_state(state)


def _set_cuda_rng_state(new_state, device=-1):  # reference name kept for API parity
    _set_device_rng_state(new_state)


def _device_manual_seed(seed):
    if use_cuda():
        torch.cuda.manual_seed(seed)
    else:
        torch.manual_seed(seed)


class CudaRNGStatesTracker:
    """Named generator states; ``fork(name)`` runs a region under that state and writes the
    advanced state back, restoring the ambient state afterwards."""

    def __init__(self):
        self.states_ = {}
        self.seeds_ = set()

    def reset(self):
        self.states_ = {}
        self.seeds_ = set()

    def get_states(self):
        return dict(self.states_)

    def set_states(self, states):
        self.states_ = states

    def add(self, name, seed):
        if seed in self.seeds_:
            raise Exception("seed {} already exists".format(seed))
        self.seeds_.add(seed)
        if name in self.states_:
            raise Exception("cuda rng state {} already exists".format(name))
        orig = _get_device_rng_state()
        _device_manual_seed(seed)
        self.states_[name] = _get_device_rng_state()
        _set_device_rng_state(orig)

    @contextlib.contextmanager
    def fork(self, name=_MODEL_PARALLEL_RNG_TRACKER_NAME):
        if name not in self.states_:
            raise Exception("cuda rng state {} is not added".format(name))
        orig = _get_device_rng_state()
        _set_device_rng_state(self.states_[name])
        try:
            yield
        finally:
            self.states_[name] = _get_device_rng_state()
            _set_device_rng_state(orig)


_CUDA_RNG_STATE_TRACKER = CudaRNGStatesTracker()


def get_cuda_rng_tracker():
    return _CUDA_RNG_STATE_TRACKER


def model_parallel_cuda_manual_seed(seed: int) -> None:
    """Default generator: same seed inside a TP group (different across DP only if the caller
    offsets it).  ``model-parallel-rng``: seed + 2718 + tp_rank, used for dropout inside
    tensor-parallel regions."""
    offset = seed + 2718
    tp_seed = offset + ps.get_tensor_model_parallel_rank()
    _CUDA_RNG_STATE_TRACKER.reset()
    _device_manual_seed(seed)
    _CUDA_RNG_STATE_TRACKER.add(_MODEL_PARALLEL_RNG_TRACKER_NAME, tp_seed)


class CheckpointFunction(torch.autograd.Function):
    """Recompute-in-backward with RNG replay (cpu + device + tracker states)."""

    @staticmethod
    def forward(ctx, run_function, distribute_saved_activations, *args):
        ctx.run_function = run_function
        ctx.distribute_saved_activations = distribute_saved_activations
        ctx.fwd_cpu_rng_state = torch.get_rng_state()
        ctx.fwd_dev_rng_state = _get_device_rng_state()
        ctx.fwd_tracker_states = get_cuda_rng_tracker().get_states()
        with torch.no_grad():
            outputs = run_function(*args)
        if distribute_saved_activations:
            ctx.input_0_shape = args[0].data.shape
            safely_set_viewless_tensor_data(
                args[0], split_tensor_into_1d_equal_chunks(args[0].data, new_buffer=True))
        ctx.save_for_backward(*[a for a in args if isinstance(a, torch.Tensor)])
        ctx.arg_is_tensor = [isinstance(a, torch.Tensor) for a in args]
        ctx.non_tensor_args = [a for a in args if not isinstance(a, torch.Tensor)]
        return outputs

    @staticmethod
    def backward(ctx, *grads):
        if not torch.autograd._is_checkpoint_valid():
            raise RuntimeError("Checkpointing is not compatible with .grad(), "
                               "please use .backward() if possible")
        saved = list(ctx.saved_tensors)
        nont = list(ctx.non_tensor_args)
        inputs = [saved.pop(0) if is_t else nont.pop(0) for is_t in ctx.arg_is_tensor]
        if ctx.distribute_saved_activations:
            safely_set_viewless_tensor_data(
                inputs[0], gather_split_1d_tensor(inputs[0].data).view(ctx.input_0_shape))

        bwd_cpu = torch.get_rng_state()
        bwd_dev = _get_device_rng_state()
        bwd_tracker = get_cuda_rng_tracker().get_states()
        torch.set_rng_state(ctx.fwd_cpu_rng_state)
        _set_device_rng_state(ctx.fwd_dev_rng_state)
        get_cuda_rng_tracker().set_states(ctx.fwd_tracker_states)

        detached = detach_variable(tuple(inputs))
        with torch.enable_grad():
            outputs = ctx.run_function(*detached)

        torch.set_rng_state(bwd_cpu)
        _set_device_rng_state(bwd_dev)
        get_cuda_rng_tracker().set_states(bwd_tracker)

        if isinstance(outputs, torch.Tensor):
            outputs = (outputs,)
        pairs = [(o, g) for o, g in zip(outputs, grads)
                 if isinstance(o, torch.Tensor) and o.requires_grad and g is not None]
        torch.autograd.backward([p[0] for p in pairs], [p[1] for p in pairs])
        in_grads = tuple(inp.grad if isinstance(inp, torch.Tensor) else None for inp in detached)
        return (None, None) + in_grads


def checkpoint(function, distribute_saved_activations, *args):
    """Checkpoint a model or part of the model (arguments as in the reference)."""
    return CheckpointFunction.apply(function, distribute_saved_activations, *args)
