"""Small numeric / tensor helpers (parity: megatron/core/utils.py:10-124)."""
from __future__ import annotations

import torch


def ensure_divisibility(numerator: int, denominator: int) -> None:
    assert numerator % denominator == 0, f"{numerator} is not divisible by {denominator}"


def divide(numerator: int, denominator: int) -> int:
    ensure_divisibility(numerator, denominator)
    return numerator // denominator


def _kernel_make_viewless_tensor(inp: torch.Tensor, requires_grad: bool) -> torch.Tensor:
    out = torch.empty((1,), dtype=inp.dtype, device=inp.device, requires_grad=requires_grad)
    out.data = inp.data
    return out


class MakeViewlessTensor(torch.autograd.Function):
    """Autograd-transparent way of dropping the ``._base`` reference of a view so the
    underlying storage can be released (pipeline output deallocation relies on it)."""

    @staticmethod
    def forward(ctx, inp, requires_grad):
        return _kernel_make_viewless_tensor(inp, requires_grad)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


def make_viewless_tensor(inp: torch.Tensor, requires_grad: bool, keep_graph: bool) -> torch.Tensor:
    if inp._base is None:
        return inp
    if keep_graph:
        return MakeViewlessTensor.apply(inp, requires_grad)
    return _kernel_make_viewless_tensor(inp, requires_grad)


def assert_viewless_tensor(tensor, extra_msg=None):
    if isinstance(tensor, (list, tuple)):
        for t in tensor:
            assert_viewless_tensor(t, extra_msg)
        return tensor
    if not isinstance(tensor, torch.Tensor):
        return tensor
    assert tensor._base is None, (
        "Ensure tensor._base is None before setting tensor.data or storing tensor to memory "
        f"buffer. Otherwise, a memory leak will occur (and likely accumulate over iterations). {extra_msg}")
    return tensor


def safely_set_viewless_tensor_data(tensor: torch.Tensor, new_data_tensor: torch.Tensor) -> None:
    assert_viewless_tensor(
        tensor, extra_msg="FYI, tensor._base has shape %s, and new_data_tensor has shape %s." % (
            "--" if tensor._base is None else tensor._base.shape, new_data_tensor.shape))
    tensor.data = new_data_tensor
