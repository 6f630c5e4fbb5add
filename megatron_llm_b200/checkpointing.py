"""Checkpoint save / load with the reference's on-disk layout.

Parity target: megatron/checkpointing.py -- names :77-140, tracker :143-214, save :243-337, base load :414-479,
``load_args_from_checkpoint`` :482-567, ``load_checkpoint`` :570-696, ``check_checkpoint_args`` :35-66.

Layout (unchanged so checkpoints carry over):
    <dir>/latest_checkpointed_iteration.txt                       (int or "release")
    <dir>/iter_{it:07d}|release/mp_rank_{tp:02d}[_{pp:03d}]/model_optim_rng.pt
    with the distributed optimizer: .../model_rng.pt  +  mp_rank_.._{dp:03d}/optim.pt
    keys: args, checkpoint_version=3.0, iteration, model | model0..k, optimizer, opt_param_scheduler, rng_state
"""
from __future__ import annotations

import contextlib
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

from .parallel import state as ps
from .parallel.random import get_cuda_rng_tracker
from .utils import print_rank_0, unwrap_model
from .utils.device import current_device, use_cuda

_CHECKPOINT_VERSION = None


def _get_args():
    from .global_vars import get_args
    return get_args()


def set_checkpoint_version(value):
    global _CHECKPOINT_VERSION
    if _CHECKPOINT_VERSION is not None:
        assert _CHECKPOINT_VERSION == value, "checkpoint versions do not match"
    _CHECKPOINT_VERSION = value


def get_checkpoint_version():
    return _CHECKPOINT_VERSION


def check_checkpoint_args(checkpoint_args):
    """Architecture / parallelism arguments of the checkpoint must match the current run."""
    args = _get_args()

    def _compare(arg_name, old_arg_name=None):
        ck = getattr(checkpoint_args, old_arg_name if old_arg_name is not None else arg_name)
        cur = getattr(args, arg_name)
        assert ck == cur, "{} value from checkpoint ({}) is not equal to the input argument value ({}).".format(
            arg_name, ck, cur)

    _compare("num_layers")
    _compare("hidden_size")
    _compare("num_attention_heads")
    if args.vocab_file:
        _compare("max_position_embeddings")
        _compare("make_vocab_size_divisible_by")
        _compare("padded_vocab_size")
        _compare("tokenizer_type")
    if args.data_parallel_random_init:
        _compare("data_parallel_random_init")
    if get_checkpoint_version() < 3.0:
        _compare("tensor_model_parallel_size", old_arg_name="model_parallel_size")
    if get_checkpoint_version() >= 3.0:
        _compare("tensor_model_parallel_size")
        _compare("pipeline_model_parallel_size")


def ensure_directory_exists(filename):
    os.makedirs(os.path.dirname(filename), exist_ok=True)


def get_checkpoint_names(checkpoints_path, iteration, use_distributed_optimizer, release=False,
                         pipeline_parallel=None, tensor_rank=None, pipeline_rank=None):
    """(model file, optimizer file); identical when the distributed optimizer is off."""
    directory = "release" if release else "iter_{:07d}".format(iteration)
    if pipeline_parallel is None:
        pipeline_parallel = ps.get_pipeline_model_parallel_world_size() > 1
    if tensor_rank is None:
        tensor_rank = ps.get_tensor_model_parallel_rank()
    if pipeline_rank is None:
        pipeline_rank = ps.get_pipeline_model_parallel_rank()
    if not pipeline_parallel:
        common = os.path.join(checkpoints_path, directory, f"mp_rank_{tensor_rank:02d}")
    else:
        common = os.path.join(checkpoints_path, directory, f"mp_rank_{tensor_rank:02d}_{pipeline_rank:03d}")
    if use_distributed_optimizer:
        model_name = os.path.join(common, "model_rng.pt")
        optim_name = os.path.join(common + "_%03d" % ps.get_data_parallel_rank(), "optim.pt")
    else:
        model_name = optim_name = os.path.join(common, "model_optim_rng.pt")
    return model_name, optim_name


def get_checkpoint_name(checkpoints_path, iteration, release=False, pipeline_parallel=None, tensor_rank=None,
                        pipeline_rank=None):
    """This rank's single-file checkpoint path (reference checkpointing.py:77-104; the tools use it)."""
    return get_checkpoint_names(checkpoints_path, iteration, False, release, pipeline_parallel, tensor_rank,
                                pipeline_rank)[0]


def find_checkpoint_rank_0(checkpoints_path, iteration, use_distributed_optimizer, release=False):
    """Locate rank (tp=0, pp=0)'s file whether or not the checkpoint was written with PP>1."""
    for pipeline_parallel in (False, True):
        names = get_checkpoint_names(checkpoints_path, iteration, use_distributed_optimizer, release,
                                     pipeline_parallel=pipeline_parallel, tensor_rank=0, pipeline_rank=0)
        if os.path.isfile(names[0]):
            return names
    return None, None


def get_checkpoint_tracker_filename(checkpoints_path):
    return os.path.join(checkpoints_path, "latest_checkpointed_iteration.txt")


def read_metadata(tracker_filename):
    iteration, release = 0, False
    with open(tracker_filename, "r") as f:
        metastring = f.read().strip()
        try:
            iteration = int(metastring)
        except ValueError:
            release = metastring == "release"
            if not release:
                print_rank_0("ERROR: Invalid metadata file {}. Exiting".format(tracker_filename))
                sys.exit()
    assert iteration > 0 or release, "error parsing metadata file {}".format(tracker_filename)
    if dist.is_initialized():
        iters = torch.tensor([iteration], dtype=torch.long, device=current_device())
        dist.all_reduce(iters, op=dist.ReduceOp.MAX)
        max_iter = iters[0].item()
        if iteration != max_iter:
            print("WARNING: on rank {} found iteration {} in the metadata while max iteration across the ranks is "
                  "{}, replacing it with max iteration.".format(dist.get_rank(), iteration, max_iter), flush=True)
    else:
        max_iter = iteration
    return max_iter, release


def get_rng_state():
    """One entry per DP rank (gathered when --data_parallel_random_init), 5 generator states each."""
    args = _get_args()
    rng_state = {"random_rng_state": random.getstate(), "np_rng_state": np.random.get_state(),
                 "torch_rng_state": torch.get_rng_state(),
                 "cuda_rng_state": torch.cuda.get_rng_state() if use_cuda() else None,
                 "rng_tracker_states": get_cuda_rng_tracker().get_states()}
    if dist.is_initialized() and ps.get_data_parallel_world_size() > 1 and args.data_parallel_random_init:
        rng_state_list = [None for _ in range(ps.get_data_parallel_world_size())]
        dist.all_gather_object(rng_state_list, rng_state, group=ps.get_data_parallel_group())
    else:
        rng_state_list = [rng_state]
    return rng_state_list


def save_checkpoint(iteration, model, optimizer, opt_param_scheduler):
    args = _get_args()
    model = unwrap_model(model)
    if not isinstance(model, list):
        model = [model]
    release = iteration == "release"        # tools (converters, checkpoint_util) save "release" checkpoints
    where = "release checkpoint" if release else "checkpoint at iteration {:7d}".format(iteration)
    print_rank_0("saving {} to {}".format(where, args.save))
    rng_state = get_rng_state()
    model_name, optim_name = get_checkpoint_names(args.save, iteration, args.use_distributed_optimizer,
                                                  release=release)

    save_optim = optimizer is not None and not args.no_save_optim
    model_state, optim_state = {}, {}
    if not dist.is_initialized() or ps.get_data_parallel_rank() == 0:
        model_state["args"] = args
        model_state["checkpoint_version"] = 3.0
        model_state["iteration"] = iteration
        if len(model) == 1:
            model_state["model"] = model[0].state_dict_for_save_checkpoint()
        else:
            for i in range(len(model)):
                ps.set_virtual_pipeline_model_parallel_rank(i)
                model_state["model%d" % i] = model[i].state_dict_for_save_checkpoint()
        if not args.no_save_rng:
            model_state["rng_state"] = rng_state
    if save_optim and (args.use_distributed_optimizer or not dist.is_initialized()
                       or ps.get_data_parallel_rank() == 0):
        optim_state["optimizer"] = optimizer.state_dict()
        if opt_param_scheduler is not None:
            optim_state["opt_param_scheduler"] = opt_param_scheduler.state_dict()
    elif (not dist.is_initialized() or ps.get_data_parallel_rank() == 0) and opt_param_scheduler is not None \
            and not args.no_save_optim:
        optim_state["opt_param_scheduler"] = opt_param_scheduler.state_dict()

    if args.use_distributed_optimizer:
        if model_state:
            ensure_directory_exists(model_name)
            torch.save(model_state, model_name)
        if optim_state:
            ensure_directory_exists(optim_name)
            torch.save(optim_state, optim_name)
    else:
        state = {**model_state, **optim_state}
        if state:
            ensure_directory_exists(model_name)
            torch.save(state, model_name)

    if dist.is_initialized():
        dist.barrier()
    print_rank_0("  successfully saved {} to {}".format(where, args.save))
    if not dist.is_initialized() or dist.get_rank() == 0:
        with open(get_checkpoint_tracker_filename(args.save), "w") as f:
            f.write(str(iteration))
    if dist.is_initialized():
        dist.barrier()


def _transpose_first_dim(t, num_splits, num_splits_first, model):
    input_shape = t.size()
    while hasattr(model, "module"):
        model = model.module
    attention_module = model.language_model.encoder.layers[0].self_attention
    hn = attention_module.hidden_size_per_attention_head
    np_ = attention_module.num_attention_heads_per_partition
    if num_splits_first:
        # [num_splits * np * hn, h] -> [np * num_splits * hn, h]
        t = t.view(num_splits, np_, hn, *input_shape[1:]).transpose(0, 1).contiguous()
    else:
        # [np * hn * num_splits, h] -> [np * num_splits * hn, h]
        t = t.view(np_, hn, num_splits, *input_shape[1:]).transpose(1, 2).contiguous()
    return t.view(*input_shape)


def fix_query_key_value_ordering(model, checkpoint_version):
    """Checkpoints older than v2.0 store QKV in a different row order (MHA only; skipped for GQA/MQA)."""
    if checkpoint_version is None or checkpoint_version >= 2.0:
        return
    if isinstance(model, list):
        assert len(model) == 1
        model = model[0]
    args = _get_args()
    if args.num_attention_heads_kv != args.num_attention_heads:
        return
    for name, param in model.named_parameters():
        if name.endswith((".query_key_value.weight", ".query_key_value.bias")):
            fixed = _transpose_first_dim(param.data, 3, checkpoint_version == 0, model) if checkpoint_version in (0, 1.0) \
                else None
            if fixed is None:
                print_rank_0(f"Invalid checkpoint version {checkpoint_version}.")
                sys.exit()
            param.data.copy_(fixed)
        if name.endswith((".key_value.weight", ".key_value.bias")):
            fixed = _transpose_first_dim(param.data, 2, checkpoint_version == 0, model)
            param.data.copy_(fixed)
    print_rank_0(" succesfully fixed query-key-values ordering for checkpoint version {}".format(checkpoint_version))


def _torch_load(path):
    # checkpoints pickle the argparse Namespace -> weights_only must be off (torch>=2.6 defaults it on)
    return torch.load(path, map_location="cpu", weights_only=False)


@contextlib.contextmanager
def legacy_pickle_modules():
    """Module paths old checkpoints were pickled under: ``megatron.*`` (the reference; aliased onto this package by the
    ``megatron`` compat package) and ``fp16.loss_scaler`` (its ancestors' loss scaler, reference checkpointing.py:
    462-474).  The unpickler imports the parent package too, so a stub ``fp16`` is registered for the duration."""
    import types
    import megatron  # noqa: F401  (compat alias package at the repo root)
    from .fp16_deprecated import loss_scaler
    added = []
    if "fp16.loss_scaler" not in sys.modules:
        parent = types.ModuleType("fp16")
        parent.loss_scaler = loss_scaler
        sys.modules["fp16"], sys.modules["fp16.loss_scaler"] = parent, loss_scaler
        added = ["fp16", "fp16.loss_scaler"]
    try:
        yield
    finally:
        for name in added:
            sys.modules.pop(name, None)


def _load_base_checkpoint(load_dir, use_distributed_optimizer, rank0=False, iteration=None):
    """Returns (model_state_dict, optim_state_dict, release)."""
    tracker = get_checkpoint_tracker_filename(load_dir)
    if not os.path.isfile(tracker):
        if not rank0:
            print_rank_0("WARNING: could not find the metadata file {} ".format(tracker))
            print_rank_0("    will not load any checkpoints and will start from random")
        return None, None, False
    if iteration is None:
        iteration, release = read_metadata(tracker)
    else:
        release = False
    if rank0:
        names = find_checkpoint_rank_0(load_dir, iteration, use_distributed_optimizer, release)
    else:
        names = get_checkpoint_names(load_dir, iteration, use_distributed_optimizer, release)
        if release:
            print_rank_0(f" loading release checkpoint from {load_dir}")
        else:
            print_rank_0(f" loading checkpoint from {load_dir} at iteration {iteration}")
    model_name, optim_name = names
    if use_distributed_optimizer and not os.path.isfile(model_name):
        # a checkpoint written without the distributed optimizer (e.g. converted HF weights or a resharded checkpoint)
        # keeps everything in model_optim_rng.pt; the model part of it is loadable (relaxes the reference, which fails)
        alt = os.path.join(os.path.dirname(model_name), "model_optim_rng.pt")
        if os.path.isfile(alt):
            model_name = alt
    try:
        model_state = _torch_load(model_name)
        optim_state = _torch_load(optim_name) if (use_distributed_optimizer and os.path.isfile(optim_name)) \
            else model_state
    except ModuleNotFoundError:
        with legacy_pickle_modules():
            model_state = _torch_load(model_name)
            optim_state = _torch_load(optim_name) if (use_distributed_optimizer and os.path.isfile(optim_name)) \
                else model_state
    except BaseException as e:
        print_rank_0("could not load the checkpoint")
        print_rank_0(e)
        sys.exit()
    return model_state, optim_state, release


def load_args_from_checkpoint(args, load_arg="load"):
    """Override architecture arguments with the ones stored in the checkpoint (``--use_checkpoint_args``)."""
    load_dir = getattr(args, load_arg)
    if load_dir is None:
        print("No load directory specified, using provided arguments.")
        return args
    model_state, _, _ = _load_base_checkpoint(load_dir, use_distributed_optimizer=False, rank0=True,
                                              iteration=getattr(args, "load_iters", None))
    if not model_state:
        print("Checkpoint not found to provide arguments, using provided arguments.")
        return args
    if "args" not in model_state:
        print("Checkpoint provided does not have arguments saved, using provided arguments.")
        return args
    ck = model_state["args"]
    args.iteration = model_state["iteration"]

    def _set_arg(name, old_name=None, force=False):
        if not force and getattr(args, name, None) is not None:
            return
        v = getattr(ck, old_name, None) if old_name is not None else getattr(ck, name, None)
        if v is not None:
            print(f"Setting {name} to {v} from checkpoint")
            setattr(args, name, v)

    for a in ("num_layers", "hidden_size", "ffn_hidden_size", "seq_length", "num_attention_heads",
              "num_attention_heads_kv", "kv_channels", "max_position_embeddings", "tokenizer_type", "glu_activation",
              "layernorm_epsilon", "rope_theta", "rope_scaling_factor", "sliding_window_size", "use_bias"):
        _set_arg(a)
    for a in ("padded_vocab_size", "position_embedding_type", "parallel_attn", "parallel_layernorm", "use_rms_norm",
              "tie_embed_logits", "make_vocab_size_divisible_by", "use_post_ln"):
        _set_arg(a, force=True)
    if model_state.get("checkpoint_version", 0) < 3.0:
        _set_arg("tensor_model_parallel_size", "model_parallel_size")
    else:
        _set_arg("tensor_model_parallel_size", force=True)
        _set_arg("pipeline_model_parallel_size", force=True)
        _set_arg("num_layers_per_virtual_pipeline_stage")
    return args


def load_checkpoint(model, optimizer, opt_param_scheduler, load_arg="load", strict=True):
    """Load a model checkpoint and return the iteration."""
    args = _get_args()
    load_dir = getattr(args, load_arg)
    model = unwrap_model(model)
    if not isinstance(model, list):
        model = [model]
    model_state, optim_state, release = _load_base_checkpoint(load_dir, args.use_distributed_optimizer, rank0=False,
                                                              iteration=getattr(args, "load_iters", None))
    if model_state is None:
        return 0
    set_checkpoint_version(model_state.get("checkpoint_version", 0))

    if args.finetune or release:
        iteration = 0
    else:
        try:
            iteration = model_state["iteration"]
        except KeyError:
            try:
                iteration = model_state["total_iters"]
            except KeyError:
                print_rank_0("A metadata file exists but unable to load iteration from checkpoint, exiting")
                sys.exit()

    assert args.consumed_train_samples == 0
    assert args.consumed_valid_samples == 0
    if "args" in model_state and not args.finetune:
        ck = model_state["args"]
        check_checkpoint_args(ck)
        args.consumed_train_samples = getattr(ck, "consumed_train_samples", 0)
        from .global_vars import update_num_microbatches
        update_num_microbatches(consumed_samples=args.consumed_train_samples)
        args.consumed_valid_samples = getattr(ck, "consumed_valid_samples", 0)
    else:
        print_rank_0("could not find arguments in the checkpoint ...")

    if len(model) == 1:
        model[0].load_state_dict(model_state["model"], strict=strict)
    else:
        for i in range(len(model)):
            ps.set_virtual_pipeline_model_parallel_rank(i)
            model[i].load_state_dict(model_state["model%d" % i], strict=strict)
    checkpoint_version = get_checkpoint_version()
    print_rank_0(f" checkpoint version {checkpoint_version}")
    fix_query_key_value_ordering(model, checkpoint_version)

    if not release and not args.finetune and not args.no_load_optim:
        try:
            if optimizer is not None:
                optimizer.load_state_dict(optim_state["optimizer"])
            if opt_param_scheduler is not None:
                if "lr_scheduler" in optim_state:
                    opt_param_scheduler.load_state_dict(optim_state["lr_scheduler"])
                else:
                    opt_param_scheduler.load_state_dict(optim_state["opt_param_scheduler"])
        except KeyError:
            print_rank_0("Unable to load optimizer from checkpoint. Specify --no_load_optim or --finetune to prevent "
                         "attempting to load the optimizer state, exiting ...")
            sys.exit()
    else:
        if (args.fp16 or args.bf16) and optimizer is not None:
            optimizer.reload_model_params()

    if not release and not args.finetune and not args.no_load_rng:
        try:
            if "rng_state" in model_state:
                rng_state = model_state["rng_state"][ps.get_data_parallel_rank()] \
                    if args.data_parallel_random_init else model_state["rng_state"][0]
                random.setstate(rng_state["random_rng_state"])
                np.random.set_state(rng_state["np_rng_state"])
                torch.set_rng_state(rng_state["torch_rng_state"])
                if use_cuda() and rng_state.get("cuda_rng_state") is not None:
                    torch.cuda.set_rng_state(rng_state["cuda_rng_state"])
                if not rng_state["rng_tracker_states"]:
                    raise KeyError
                get_cuda_rng_tracker().set_states(rng_state["rng_tracker_states"])
            else:
                random.setstate(model_state["random_rng_state"])
                np.random.set_state(model_state["np_rng_state"])
                torch.set_rng_state(model_state["torch_rng_state"])
                if use_cuda():
                    torch.cuda.set_rng_state(model_state["cuda_rng_state"])
                if not model_state["rng_tracker_states"]:
                    raise KeyError
                get_cuda_rng_tracker().set_states(model_state["rng_tracker_states"])
        except KeyError:
            print_rank_0("Unable to load rng state from checkpoint. Specify --no_load_rng or --finetune to prevent "
                         "attempting to load the rng state, exiting ...")
            sys.exit()

    # re-home the loaded fp16/bf16 weights into the optimizer's master copy if it did not load its own
    if dist.is_initialized():
        dist.barrier()
    print_rank_0(f"  successfully loaded checkpoint from {load_dir} at iteration {iteration}")
    return iteration


def load_biencoder_checkpoint(model, only_query_model=False, only_context_model=False, custom_load_path=None):
    """Selectively load the query / context towers of an ICT/REALM bi-encoder checkpoint."""
    args = _get_args()
    model = unwrap_model(model)
    load_path = custom_load_path if custom_load_path is not None else args.load
    tracker = get_checkpoint_tracker_filename(load_path)
    with open(tracker, "r") as f:
        iteration = int(f.read().strip())
    checkpoint_name, _ = get_checkpoint_names(load_path, iteration, args.use_distributed_optimizer, False)
    if ps.get_data_parallel_rank() == 0:
        print("global rank {} is loading checkpoint {}".format(dist.get_rank(), checkpoint_name))
    state_dict = _torch_load(checkpoint_name)
    ret_state_dict = state_dict["model"]
    if only_query_model:
        ret_state_dict.pop("context_model")
    if only_context_model:
        ret_state_dict.pop("query_model")
    assert len(model) == 1
    model[0].load_state_dict(ret_state_dict)
    dist.barrier()
    if ps.get_data_parallel_rank() == 0:
        print(" successfully loaded {}".format(checkpoint_name))
    return model
