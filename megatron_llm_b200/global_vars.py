"""Process-wide singletons: args, micro-batch calculator, tokenizer, TB/W&B writer, timers, counters,
signal handler (parity: megatron/global_vars.py:89-207)."""
from __future__ import annotations

import os
import sys
from collections import defaultdict

import torch

from . import dist_signal_handler
from .microbatches import build_num_microbatches_calculator
from .timers import Timers

_GLOBAL_ARGS = None
_GLOBAL_NUM_MICROBATCHES_CALCULATOR = None
_GLOBAL_TOKENIZER = None
_GLOBAL_TENSORBOARD_WRITER = None
_GLOBAL_ADLR_AUTORESUME = None
_GLOBAL_TIMERS = None
_GLOBAL_SIGNAL_HANDLER = None
_GLOBAL_COUNTERS = None


def _ensure_var_is_initialized(var, name):
    assert var is not None, "{} is not initialized.".format(name)


def _ensure_var_is_not_initialized(var, name):
    assert var is None, "{} is already initialized.".format(name)


def get_args():
    _ensure_var_is_initialized(_GLOBAL_ARGS, "args")
    return _GLOBAL_ARGS


def args_initialized() -> bool:
    return _GLOBAL_ARGS is not None


def get_num_microbatches():
    return _GLOBAL_NUM_MICROBATCHES_CALCULATOR.get()


def get_current_global_batch_size():
    return _GLOBAL_NUM_MICROBATCHES_CALCULATOR.get_current_global_batch_size()


def update_num_microbatches(consumed_samples, consistency_check=True):
    _GLOBAL_NUM_MICROBATCHES_CALCULATOR.update(consumed_samples, consistency_check)


def get_tokenizer():
    _ensure_var_is_initialized(_GLOBAL_TOKENIZER, "tokenizer")
    return _GLOBAL_TOKENIZER


def get_tensorboard_writer():
    return _GLOBAL_TENSORBOARD_WRITER


def get_adlr_autoresume():
    return _GLOBAL_ADLR_AUTORESUME


def get_timers():
    _ensure_var_is_initialized(_GLOBAL_TIMERS, "timers")
    return _GLOBAL_TIMERS


def get_counters():
    _ensure_var_is_initialized(_GLOBAL_COUNTERS, "counters")
    return _GLOBAL_COUNTERS


def get_signal_handler():
    _ensure_var_is_initialized(_GLOBAL_SIGNAL_HANDLER, "signal handler")
    return _GLOBAL_SIGNAL_HANDLER


def _set_signal_handler():
    global _GLOBAL_SIGNAL_HANDLER
    _ensure_var_is_not_initialized(_GLOBAL_SIGNAL_HANDLER, "signal handler")
    _GLOBAL_SIGNAL_HANDLER = dist_signal_handler.DistributedSignalHandler().__enter__()


def set_global_variables(args):
    """Set args, tokenizer, tensorboard-writer, adlr-autoresume, timers, counters."""
    global _GLOBAL_ARGS, _GLOBAL_NUM_MICROBATCHES_CALCULATOR, _GLOBAL_COUNTERS
    assert args is not None
    _ensure_var_is_not_initialized(_GLOBAL_ARGS, "args")
    _GLOBAL_ARGS = args
    _GLOBAL_NUM_MICROBATCHES_CALCULATOR = build_num_microbatches_calculator(args)
    _GLOBAL_COUNTERS = defaultdict(int)
    if args.vocab_file or args.tokenizer_type is not None:
        _build_tokenizer(args)
    _set_tensorboard_writer(args)
    _set_adlr_autoresume(args)
    _set_timers(args)
    if args.exit_signal_handler:
        _set_signal_handler()


def unset_global_variables():
    """Tear down every singleton (tests / tools that re-initialise in one process)."""
    global _GLOBAL_ARGS, _GLOBAL_NUM_MICROBATCHES_CALCULATOR, _GLOBAL_TOKENIZER, _GLOBAL_TENSORBOARD_WRITER
    global _GLOBAL_ADLR_AUTORESUME, _GLOBAL_TIMERS, _GLOBAL_SIGNAL_HANDLER, _GLOBAL_COUNTERS
    _GLOBAL_ARGS = _GLOBAL_NUM_MICROBATCHES_CALCULATOR = _GLOBAL_TOKENIZER = None
    _GLOBAL_TENSORBOARD_WRITER = _GLOBAL_ADLR_AUTORESUME = _GLOBAL_TIMERS = None
    _GLOBAL_SIGNAL_HANDLER = _GLOBAL_COUNTERS = None


def set_args(args):
    global _GLOBAL_ARGS
    _GLOBAL_ARGS = args


def _build_tokenizer(args):
    global _GLOBAL_TOKENIZER
    _ensure_var_is_not_initialized(_GLOBAL_TOKENIZER, "tokenizer")
    from .tokenizer import build_tokenizer
    _GLOBAL_TOKENIZER = build_tokenizer(args)
    return _GLOBAL_TOKENIZER


def rebuild_tokenizer(args):
    global _GLOBAL_TOKENIZER
    _GLOBAL_TOKENIZER = None
    return _build_tokenizer(args)


def _set_tensorboard_writer(args):
    """TensorBoard (or the W&B shim that mirrors into it) lives on the LAST rank only."""
    global _GLOBAL_TENSORBOARD_WRITER
    _ensure_var_is_not_initialized(_GLOBAL_TENSORBOARD_WRITER, "tensorboard writer")
    if getattr(args, "wandb_logger", False) and args.rank == (args.world_size - 1):
        from .wandb_logger import WandBConfig, WandbTBShim
        _GLOBAL_TENSORBOARD_WRITER = WandbTBShim(WandBConfig.from_args(args))
        return
    if getattr(args, "tensorboard_dir", None) and args.rank == (args.world_size - 1):
        try:
            from torch.utils.tensorboard import SummaryWriter
            print("> setting tensorboard ...")
            _GLOBAL_TENSORBOARD_WRITER = SummaryWriter(log_dir=args.tensorboard_dir,
                                                       max_queue=args.tensorboard_queue_size)
        except ModuleNotFoundError:
            print("WARNING: TensorBoard writing requested but is not available, no TensorBoard logs will be "
                  "written.", flush=True)


def _set_adlr_autoresume(args):
    global _GLOBAL_ADLR_AUTORESUME
    _ensure_var_is_not_initialized(_GLOBAL_ADLR_AUTORESUME, "adlr autoresume")
    if args.adlr_autoresume:
        if args.rank == 0:
            print("enabling autoresume ...", flush=True)
        sys.path.append(os.environ.get("SUBMIT_SCRIPTS", "."))
        try:
            from userlib.auto_resume import AutoResume
        except BaseException:
            print("ADLR autoresume is not available, exiting ...")
            sys.exit()
        _GLOBAL_ADLR_AUTORESUME = AutoResume


def _set_timers(args):
    global _GLOBAL_TIMERS
    _ensure_var_is_not_initialized(_GLOBAL_TIMERS, "timers")
    _GLOBAL_TIMERS = Timers(args.timing_log_level, args.timing_log_option)
