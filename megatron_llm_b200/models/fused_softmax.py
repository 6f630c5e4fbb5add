"""scale + mask + softmax dispatcher (parity: megatron/model/fused_softmax.py:90-213).

The sm_100a kernels (csrc/softmax.cu) have no sequence-length envelope, so ``is_kernel_available`` only checks
dtype/device, unlike the reference's ``16 < sk <= 4096`` / ``sq % 4`` / ``b*np % 4`` constraints (:152-172).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .enums import AttnMaskType


class ScaledUpperTriangMaskedSoftmax:
    @staticmethod
    def apply(inputs, scale):
        return ops.scaled_upper_triang_masked_softmax(inputs, scale)


class ScaledMaskedSoftmax:
    @staticmethod
    def apply(inputs, mask, scale):
        return ops.scaled_masked_softmax(inputs, mask, scale)


class ScaledSoftmax:
    @staticmethod
    def apply(inputs, scale):
        return ops.scaled_softmax(inputs, scale)


class FusedScaleMaskSoftmax(nn.Module):
    """input [b, np, sq, sk] -> softmax(scale * input with mask)."""

    def __init__(self, input_in_fp16, input_in_bf16, attn_mask_type, scaled_masked_softmax_fusion, mask_func,
                 softmax_in_fp32, scale):
        super().__init__()
        self.input_in_fp16, self.input_in_bf16 = input_in_fp16, input_in_bf16
        assert not (input_in_fp16 and input_in_bf16), "both fp16 and bf16 flags cannot be active at the same time."
        self.input_in_float16 = input_in_fp16 or input_in_bf16
        self.attn_mask_type = attn_mask_type
        self.scaled_masked_softmax_fusion = scaled_masked_softmax_fusion
        self.mask_func = mask_func
        self.softmax_in_fp32 = softmax_in_fp32
        self.scale = scale
        assert self.scale is None or softmax_in_fp32, "softmax should be in fp32 when scaled"

    def forward(self, input, mask):
        assert input.dim() == 4
        if self.is_kernel_available(mask, *input.size()):
            return self.forward_fused_softmax(input, mask)
        return self.forward_torch_softmax(input, mask)

    def is_kernel_available(self, mask, b, np, sq, sk):
        return bool(self.scaled_masked_softmax_fusion and self.input_in_float16)

    @staticmethod
    def get_batch_per_block(sq, sk, b, np):
        """Softmax rows handled by one thread block.  The reference's warp-per-row kernels pack several rows per
        block and need ``sq * b * np`` to divide by this (fused_softmax.py:155-166, :209-213); csrc/softmax.cu runs
        one block per row, so any shape qualifies."""
        return 1

    def forward_fused_softmax(self, input, mask):
        scale = self.scale if self.scale is not None else 1.0
        if self.attn_mask_type == AttnMaskType.causal and input.size(2) == input.size(3):
            return ops.scaled_upper_triang_masked_softmax(input, scale)
        if mask is not None:
            return ops.scaled_masked_softmax(input, mask, scale)
        return ops.scaled_softmax(input, scale)

    def forward_torch_softmax(self, input, mask):
        if self.input_in_float16 and self.softmax_in_fp32:
            input = input.float()
        if self.scale is not None:
            input = input * self.scale
        mask_output = self.mask_func(input, mask) if mask is not None else input
        probs = torch.nn.Softmax(dim=-1)(mask_output)
        if self.input_in_float16 and self.softmax_in_fp32:
            probs = probs.half() if self.input_in_fp16 else probs.bfloat16()
        return probs
