"""Blackwell-native (sm_100a) 3D-parallel LLM training engine with the capabilities, public API and checkpoint
layout of epfLLM/Megatron-LLM.  The compat package ``megatron`` at the repository root maps the reference's
import paths (``megatron.core.mpu``, ``megatron.model`` ...) onto these modules.

Public surface (parity: megatron/__init__.py:4-18)."""
from .global_vars import (get_args, get_adlr_autoresume, get_counters, get_current_global_batch_size,
                          get_num_microbatches, get_signal_handler, get_tensorboard_writer, get_timers,
                          get_tokenizer, update_num_microbatches)
from .utils import is_last_rank, print_all_nodes, print_rank_0, print_rank_last

__version__ = "0.1.0"
