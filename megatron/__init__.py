"""Compatibility package: the reference's import surface (``megatron.*``) mapped onto ``megatron_llm_b200``.

A user script written against epfLLM/Megatron-LLM (``from megatron import get_args``, ``from megatron.core import
mpu, tensor_parallel``, ``from megatron.model import LlamaModel`` ...) runs unchanged.  Modules are aliased
lazily through a meta-path finder, so ``sys.modules['megatron.core.parallel_state']`` IS
``megatron_llm_b200.parallel.state`` (no duplicated state).  Checkpoints written by the reference pickle
``megatron.*`` class paths (e.g. the args Namespace holds ``megatron.model.enums.PositionEmbeddingType``); they
unpickle through these aliases too.
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

_T = "megatron_llm_b200"

_ALIASES = {
    "megatron.core": f"{_T}.core_compat",
    "megatron.core.parallel_state": f"{_T}.parallel.state",
    "megatron.core.mpu": f"{_T}.parallel.state",
    "megatron.mpu": f"{_T}.parallel.state",
    "megatron.core.utils": f"{_T}.utils.core_utils",
    "megatron.core.tensor_parallel": f"{_T}.parallel",
    "megatron.core.tensor_parallel.layers": f"{_T}.parallel.layers",
    "megatron.core.tensor_parallel.mappings": f"{_T}.parallel.mappings",
    "megatron.core.tensor_parallel.cross_entropy": f"{_T}.parallel.cross_entropy",
    "megatron.core.tensor_parallel.random": f"{_T}.parallel.random",
    "megatron.core.tensor_parallel.data": f"{_T}.parallel.data",
    "megatron.core.tensor_parallel.utils": f"{_T}.parallel.tp_utils",
    "megatron.model": f"{_T}.models",
    "megatron.model.distributed": f"{_T}.parallel.ddp",
    "megatron.model.fused_layer_norm": f"{_T}.models.norms",
    "megatron.model.fused_bias_gelu": f"{_T}.models.activations",
    "megatron.model.glu_activations": f"{_T}.models.activations",
    "megatron.model.utils": f"{_T}.models.activations",
    "megatron.p2p_communication": f"{_T}.parallel.p2p",
    "megatron.schedules": f"{_T}.parallel.schedules",
    "megatron.fused_kernels": f"{_T}.ops",
    # very old checkpoints pickle their loss scaler under these paths (reference checkpointing.py:462-474)
    "megatron.fp16_deprecated": f"{_T}.fp16_deprecated",
    "megatron.fp16_deprecated.loss_scaler": f"{_T}.fp16_deprecated.loss_scaler",
    "megatron.fp16": f"{_T}.fp16_deprecated",
    "megatron.fp16.loss_scaler": f"{_T}.fp16_deprecated.loss_scaler",
}
# modules whose name is identical below the package root
for _name in ("arguments", "checkpointing", "dist_signal_handler", "global_vars", "initialize", "memory", "microbatches",
              "optimizer_param_scheduler", "timers", "training", "utils", "wandb_logger", "metrics", "indexer",
              "text_generation_server", "optimizer", "optimizer.optimizer", "optimizer.distrib_optimizer",
              "optimizer.clip_grads", "optimizer.grad_scaler", "data", "tokenizer", "tokenizer.tokenizer",
              "tokenizer.bert_tokenization", "tokenizer.gpt2_tokenization", "text_generation"):
    _ALIASES.setdefault(f"megatron.{_name}", f"{_T}.{_name}")
for _name in ("transformer", "language_model", "gpt_model", "llama_model", "falcon_model", "mistral_model", "module",
              "enums", "fused_softmax", "positional_embeddings", "bert_model", "t5_model", "biencoder_model",
              "classification", "multiple_choice"):
    _ALIASES.setdefault(f"megatron.model.{_name}", f"{_T}.models.{_name}")


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        mod = importlib.import_module(self.target)
        return mod

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if not fullname.startswith("megatron."):
            return None
        tgt = _ALIASES.get(fullname)
        if tgt is None:
            # generic fall-through: megatron.data.x -> megatron_llm_b200.data.x, etc.
            for prefix, repl in (("megatron.data.", f"{_T}.data."), ("megatron.text_generation.", f"{_T}.text_generation."),
                                 ("megatron.tokenizer.", f"{_T}.tokenizer."), ("megatron.optimizer.", f"{_T}.optimizer.")):
                if fullname.startswith(prefix):
                    tgt = repl + fullname[len(prefix):]
                    break
        if tgt is None:
            return None
        try:
            importlib.import_module(tgt)
        except ModuleNotFoundError:
            return None
        spec = importlib.machinery.ModuleSpec(fullname, _AliasLoader(tgt), is_package=True)
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

from megatron_llm_b200 import (get_adlr_autoresume, get_args, get_counters, get_current_global_batch_size,  # noqa: E402
                               get_num_microbatches, get_signal_handler, get_tensorboard_writer, get_timers,
                               get_tokenizer, is_last_rank, print_all_nodes, print_rank_0, print_rank_last,
                               update_num_microbatches)
from megatron_llm_b200.initialize import initialize_megatron  # noqa: E402,F401
