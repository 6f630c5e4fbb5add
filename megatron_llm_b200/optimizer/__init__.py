"""Optimizer factory (parity: megatron/optimizer/__init__.py: param groups :13-61, selection :63-144)."""
from __future__ import annotations

from .grad_scaler import ConstantGradScaler, DynamicGradScaler
from .optimizer import FP32Optimizer, Float16OptimizerWithFloat16Params, MegatronOptimizer
from .distrib_optimizer import DistributedOptimizer


def get_param_groups(modules, no_weight_decay_cond, scale_lr_cond, lr_mult):
    """Four groups (wd x lr-scale); by default biases and 1-D params (norm weights) get no weight decay."""
    wd_no_scale, wd_scale, nowd_no_scale, nowd_scale = [], [], [], []
    for module in modules:
        for name, param in module.named_parameters():
            if not param.requires_grad:
                continue
            if no_weight_decay_cond is not None:
                no_wd = no_weight_decay_cond(name, param)
            else:
                no_wd = name.endswith(".bias") or len(param.shape) == 1
            scale_lr = scale_lr_cond(name, param) if scale_lr_cond is not None else False
            if not no_wd and not scale_lr:
                wd_no_scale.append(param)
            elif not no_wd and scale_lr:
                wd_scale.append(param)
            elif no_wd and not scale_lr:
                nowd_no_scale.append(param)
            else:
                nowd_scale.append(param)
    groups = []
    for params, wd_mult, lr_m in ((wd_no_scale, 1.0, 1.0), (wd_scale, 1.0, lr_mult), (nowd_no_scale, 0.0, 1.0),
                                  (nowd_scale, 0.0, lr_mult)):
        if len(params):
            groups.append({"params": params, "wd_mult": wd_mult, "lr_mult": lr_m})
    return groups


def get_megatron_optimizer(model, no_weight_decay_cond=None, scale_lr_cond=None, lr_mult=1.0):
    from ..global_vars import get_args
    args = get_args()
    param_groups = get_param_groups(model, no_weight_decay_cond, scale_lr_cond, lr_mult)
    for g in param_groups:
        g["lr"] = args.lr * g["lr_mult"]
        g["weight_decay"] = args.weight_decay * g["wd_mult"]
    if args.optimizer == "adam":
        config = {"name": "adam", "betas": (args.adam_beta1, args.adam_beta2), "eps": args.adam_eps}
    elif args.optimizer == "sgd":
        config = {"name": "sgd", "momentum": args.sgd_momentum}
    else:
        raise Exception("{} optimizer is not supported.".format(args.optimizer))
    config["param_groups"] = param_groups

    params_have_main_grad = args.DDP_impl == "local"
    if args.fp16 or args.bf16 or args.use_distributed_optimizer:
        grad_scaler = None
        if args.loss_scale:
            grad_scaler = ConstantGradScaler(args.loss_scale)
        elif args.fp16:
            grad_scaler = DynamicGradScaler(initial_scale=args.initial_loss_scale, min_scale=args.min_loss_scale,
                                            growth_factor=2.0, backoff_factor=0.5,
                                            growth_interval=args.loss_scale_window, hysteresis=args.hysteresis)
        cls = DistributedOptimizer if args.use_distributed_optimizer else Float16OptimizerWithFloat16Params
        return cls(config, args.clip_grad, args.log_num_zeros_in_grad, params_have_main_grad,
                   args.use_contiguous_buffers_in_local_ddp, args.fp16, args.bf16, args.params_dtype, grad_scaler,
                   model)
    return FP32Optimizer(config, args.clip_grad, args.log_num_zeros_in_grad, params_have_main_grad,
                         args.use_contiguous_buffers_in_local_ddp, model)
