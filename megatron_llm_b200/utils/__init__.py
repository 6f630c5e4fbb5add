"""General utilities (parity: megatron/utils.py)."""
from __future__ import annotations

import os
import sys

import torch
import torch.distributed as dist

from .device import current_device, use_cuda


def _args():
    from ..global_vars import get_args
    return get_args()


def unwrap_model(model, module_instances=None):
    from ..parallel.ddp import DistributedDataParallel as LocalDDP
    from ..models.module import Float16Module
    from torch.nn.parallel import DistributedDataParallel as torchDDP
    if module_instances is None:
        module_instances = (torchDDP, LocalDDP, Float16Module)
    return_list = True
    if not isinstance(model, list):
        model = [model]
        return_list = False
    out = []
    for m in model:
        while isinstance(m, module_instances):
            m = m.module
        out.append(m)
    return out if return_list else out[0]


def calc_params_l2_norm(model):
    """L2 norm of the parameters, counting shared / TP-duplicated params once, reduced over the
    model-parallel group."""
    from ..parallel import state as ps
    from ..parallel.layers import param_is_not_tensor_parallel_duplicate
    args = _args()
    if not isinstance(model, list):
        model = [model]
    sq = torch.zeros(1, dtype=torch.float32, device=current_device())
    for m in model:
        for p in m.parameters():
            if getattr(p, "shared", False) or not param_is_not_tensor_parallel_duplicate(p):
                continue
            sq += p.data.float().pow(2).sum() if not args.bf16 else p.data.float().pow(2).sum()
    if dist.is_initialized() and ps.model_parallel_is_initialized():
        dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=ps.get_model_parallel_group())
    return sq.item() ** 0.5


def average_losses_across_data_parallel_group(losses):
    """Reduce a list of scalar tensors across the data-parallel group (mean)."""
    from ..parallel import state as ps
    averaged = torch.cat([l.clone().detach().view(1).float() for l in losses])
    if dist.is_initialized() and ps.model_parallel_is_initialized() and ps.get_data_parallel_world_size() > 1:
        dist.all_reduce(averaged, group=ps.get_data_parallel_group())
        averaged = averaged / ps.get_data_parallel_world_size()
    return averaged


def report_memory(name):
    if not use_cuda():
        return
    mb = 1024.0 * 1024.0
    s = name + " memory (MB)"
    s += " | allocated: {}".format(torch.cuda.memory_allocated() / mb)
    s += " | max allocated: {}".format(torch.cuda.max_memory_allocated() / mb)
    s += " | reserved: {}".format(torch.cuda.memory_reserved() / mb)
    s += " | max reserved: {}".format(torch.cuda.max_memory_reserved() / mb)
    from ..parallel import state as ps
    if ps.get_data_parallel_rank() == 0:
        print("[Rank {}] {}".format(dist.get_rank() if dist.is_initialized() else 0, s), flush=True)


def print_params_min_max_norm(optimizer, iteration):
    rank = dist.get_rank() if dist.is_initialized() else 0
    s = "iteration, rank, index, tensor-model-parallel, min, max, norm\n"
    index = 0
    for group in optimizer.param_groups:
        for p in group["params"]:
            index += 1
            s += "{:7d}, {:4d}, {:4d}, {:2d}, {:.6E}, {:.6E}, {:.6E}\n".format(
                iteration, rank, index, int(getattr(p, "tensor_model_parallel", False)),
                p.data.min().item(), p.data.max().item(), torch.linalg.norm(p.data.float()).item())
    print(s, flush=True)


def check_adlr_autoresume_termination(iteration, model, optimizer, opt_param_scheduler):
    from ..checkpointing import save_checkpoint
    from ..global_vars import get_adlr_autoresume
    autoresume = get_adlr_autoresume()
    if dist.is_initialized():
        dist.barrier()
    if autoresume.termination_requested():
        if _args().save:
            save_checkpoint(iteration, model, optimizer, opt_param_scheduler)
        print_rank_0(">>> autoresume termination request found!")
        if (dist.get_rank() if dist.is_initialized() else 0) == 0:
            autoresume.request_resume()
        print_rank_0(">>> training terminated. Returning")
        sys.exit(0)


def get_ltor_masks_and_position_ids(data, eod_token, reset_position_ids, reset_attention_mask, eod_mask_loss,
                                    build_attention_mask: bool = True):
    """Left-to-right (causal) masks and position ids.

    Returns (attention_mask [att_b,1,s,s] bool with True = masked, or None; loss_mask [b,s] float;
    position_ids [b,s] long).  ``build_attention_mask=False`` skips the O(s^2) mask that the flash /
    tcgen05 attention path never reads (the reference always materialises it: utils.py:152-154)."""
    b, s = data.size()
    att_b = b if reset_attention_mask else 1
    attention_mask = None
    if build_attention_mask or reset_attention_mask:
        attention_mask = torch.tril(torch.ones((att_b, s, s), device=data.device)).view(att_b, 1, s, s)
    loss_mask = torch.ones(data.size(), dtype=torch.float, device=data.device)
    if eod_mask_loss:
        loss_mask[data == eod_token] = 0.0
    position_ids = torch.arange(s, dtype=torch.long, device=data.device).unsqueeze(0).expand_as(data)
    if reset_position_ids:
        position_ids = position_ids.clone()
    if reset_position_ids or reset_attention_mask:
        for bi in range(b):
            eod_index = position_ids[bi, data[bi] == eod_token]
            if reset_position_ids:
                eod_index = eod_index.clone()
            prev = 0
            for j in range(eod_index.size(0)):
                i = eod_index[j]
                if reset_attention_mask:
                    attention_mask[bi, 0, (i + 1):, :(i + 1)] = 0
                if reset_position_ids:
                    position_ids[bi, (i + 1):] -= (i + 1 - prev)
                    prev = i + 1
    if attention_mask is not None:
        attention_mask = attention_mask < 0.5
    return attention_mask, loss_mask, position_ids


def print_rank_0(message):
    if dist.is_initialized():
        if dist.get_rank() == 0:
            print(message, flush=True)
    else:
        print(message, flush=True)


def is_last_rank():
    return (not dist.is_initialized()) or dist.get_rank() == (dist.get_world_size() - 1)


def print_rank_last(message):
    if is_last_rank():
        print(message, flush=True)


def is_last_local_rank():
    """Last process of this node (reference utils.py:218-219; LOCAL_WORLD_SIZE from torchrun, 1 if absent)."""
    local_rank = getattr(_args(), "local_rank", None)
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    return local_rank == int(os.environ.get("LOCAL_WORLD_SIZE", "1")) - 1


def print_all_nodes(message):
    """Print on the last local rank of every node (LOCAL_WORLD_SIZE from torchrun; 1 if absent)."""
    if dist.is_initialized():
        lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        if (dist.get_rank() + 1) % lws == 0:
            print(message, flush=True)
    else:
        print(message, flush=True)
