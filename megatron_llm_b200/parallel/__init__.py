"""Parallelism: process-group state ("mpu"), tensor/sequence-parallel layers and mappings, pipeline schedules and
P2P, data-parallel wrapper, symmetric-memory communicators and fused GEMM+collective dispatch."""
from . import state as parallel_state  # noqa: F401

# ``tensor_parallel`` namespace of the reference (megatron/core/tensor_parallel/__init__.py:41-65)
from .cross_entropy import vocab_parallel_cross_entropy, vocab_parallel_max_indices  # noqa: F401
from .data import broadcast_data  # noqa: F401
from .layers import (ColumnParallelLinear, RowParallelLinear, VocabParallelEmbedding,  # noqa: F401
                     copy_tensor_model_parallel_attributes, linear_with_grad_accumulation_and_async_allreduce,
                     param_is_not_tensor_parallel_duplicate, set_defaults_if_not_set_tensor_model_parallel_attributes,
                     set_tensor_model_parallel_attributes)
from .mappings import (copy_to_tensor_model_parallel_region, gather_from_sequence_parallel_region,  # noqa: F401
                       gather_from_tensor_model_parallel_region, reduce_from_tensor_model_parallel_region,
                       reduce_scatter_to_sequence_parallel_region, scatter_to_sequence_parallel_region,
                       scatter_to_tensor_model_parallel_region)
from .random import checkpoint, get_cuda_rng_tracker, model_parallel_cuda_manual_seed  # noqa: F401
from .tp_utils import (VocabUtility, gather_split_1d_tensor, split_tensor_along_last_dim,  # noqa: F401
                       split_tensor_into_1d_equal_chunks)
