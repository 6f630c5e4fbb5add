"""Learning-rate / weight-decay schedules (parity: megatron/optimizer_param_scheduler.py:9-228).

Steps are counted in samples (``increment = micro_batch * DP * num_microbatches``).  LR: linear warm-up then
constant / linear / cosine / inverse-square-root decay to ``min_lr``; WD: constant / linear / cosine ramp."""
from __future__ import annotations

import math

from .utils import print_rank_0


class OptimizerParamScheduler:
    def __init__(self, optimizer, max_lr, min_lr, lr_warmup_steps, lr_decay_steps, lr_decay_style, start_wd,
                 end_wd, wd_incr_steps, wd_incr_style, use_checkpoint_opt_param_scheduler=True,
                 override_opt_param_scheduler=False):
        self.optimizer = optimizer
        self.max_lr = float(max_lr)
        self.min_lr = min_lr
        assert self.min_lr >= 0.0
        assert self.max_lr >= self.min_lr
        self.lr_warmup_steps = lr_warmup_steps
        self.num_steps = 0
        self.lr_decay_steps = lr_decay_steps
        assert self.lr_decay_steps > 0
        assert self.lr_warmup_steps < self.lr_decay_steps
        self.lr_decay_style = lr_decay_style
        self.start_wd, self.end_wd = start_wd, end_wd
        assert self.start_wd >= 0.0
        assert self.end_wd >= self.start_wd
        self.wd_incr_steps = wd_incr_steps
        self.wd_incr_style = wd_incr_style
        self.override_opt_param_scheduler = override_opt_param_scheduler
        self.use_checkpoint_opt_param_scheduler = use_checkpoint_opt_param_scheduler
        if self.override_opt_param_scheduler:
            assert not self.use_checkpoint_opt_param_scheduler, "both override and use-checkpoint are set."
        self.step(0)
        print_rank_0("> learning rate decay style: {}".format(self.lr_decay_style))

    def get_wd(self):
        if self.num_steps > self.wd_incr_steps:
            return self.end_wd
        if self.wd_incr_style == "constant":
            assert self.start_wd == self.end_wd
            return self.end_wd
        ratio = float(self.num_steps) / float(self.wd_incr_steps)
        assert 0.0 <= ratio <= 1.0
        delta = self.end_wd - self.start_wd
        if self.wd_incr_style == "linear":
            coeff = ratio
        elif self.wd_incr_style == "cosine":
            coeff = 0.5 * (math.cos(math.pi * (1 - ratio)) + 1.0)
        else:
            raise Exception("{} weight decay increment style is not supported.".format(self.wd_incr_style))
        return self.start_wd + coeff * delta

    def get_lr(self):
        if self.lr_warmup_steps > 0 and self.num_steps <= self.lr_warmup_steps:
            return self.max_lr * float(self.num_steps) / float(self.lr_warmup_steps)
        if self.lr_decay_style == "constant":
            return self.max_lr
        if self.num_steps > self.lr_decay_steps:
            return self.min_lr
        if self.lr_decay_style == "inverse-square-root":
            warmup = max(self.lr_warmup_steps, 1)
            lr = self.max_lr * warmup ** 0.5 / (max(self.num_steps, 1) ** 0.5)
            return max(self.min_lr, lr)
        num_steps_ = self.num_steps - self.lr_warmup_steps
        decay_steps_ = self.lr_decay_steps - self.lr_warmup_steps
        ratio = float(num_steps_) / float(decay_steps_)
        assert 0.0 <= ratio <= 1.0
        delta = self.max_lr - self.min_lr
        if self.lr_decay_style == "linear":
            coeff = 1.0 - ratio
        elif self.lr_decay_style == "cosine":
            coeff = 0.5 * (math.cos(math.pi * ratio) + 1.0)
        else:
            raise Exception("{} decay style is not supported.".format(self.lr_decay_style))
        return self.min_lr + coeff * delta

    def step(self, increment):
        self.num_steps += increment
        new_lr, new_wd = self.get_lr(), self.get_wd()
        for group in self.optimizer.param_groups:
            group["lr"] = new_lr * group.get("lr_mult", 1.0)
            group["weight_decay"] = new_wd * group.get("wd_mult", 1.0)

    def state_dict(self):
        return {"max_lr": self.max_lr, "lr_warmup_steps": self.lr_warmup_steps, "num_steps": self.num_steps,
                "lr_decay_style": self.lr_decay_style, "lr_decay_steps": self.lr_decay_steps, "min_lr": self.min_lr,
                "start_wd": self.start_wd, "end_wd": self.end_wd, "wd_incr_style": self.wd_incr_style,
                "wd_incr_steps": self.wd_incr_steps}

    def _check_and_set(self, cls_value, sd_value, name):
        if self.override_opt_param_scheduler:
            print_rank_0(" > overriding {} value to {}".format(name, cls_value))
            return cls_value
        if not self.use_checkpoint_opt_param_scheduler:
            assert cls_value == sd_value, \
                f"OptimizerParamScheduler: class input value {cls_value} and checkpoint value {sd_value} for {name} do not match"
        print_rank_0(" > using checkpoint value {} for {}".format(sd_value, name))
        return sd_value

    def load_state_dict(self, sd):
        max_lr_ = sd["start_lr"] if "start_lr" in sd else sd["max_lr"]
        self.max_lr = self._check_and_set(self.max_lr, max_lr_, "learning rate")
        self.min_lr = self._check_and_set(self.min_lr, sd["min_lr"], "minimum learning rate")
        if "warmup_iter" in sd:
            warm = sd["warmup_iter"]
        elif "warmup_steps" in sd:
            warm = sd["warmup_steps"]
        else:
            warm = sd["lr_warmup_steps"]
        self.lr_warmup_steps = self._check_and_set(self.lr_warmup_steps, warm, "warmup iterations")
        if "end_iter" in sd:
            decay = sd["end_iter"]
        elif "decay_steps" in sd:
            decay = sd["decay_steps"]
        else:
            decay = sd["lr_decay_steps"]
        self.lr_decay_steps = self._check_and_set(self.lr_decay_steps, decay, "total number of iterations")
        style = sd["decay_style"] if "decay_style" in sd else sd["lr_decay_style"]
        self.lr_decay_style = self._check_and_set(self.lr_decay_style, style, "learning rate decay style")
        num_steps = sd["num_iters"] if "num_iters" in sd else sd["num_steps"]
        self.step(increment=num_steps)
        if "start_wd" in sd:
            self.start_wd = self._check_and_set(self.start_wd, sd["start_wd"], "start weight decay")
            self.end_wd = self._check_and_set(self.end_wd, sd["end_wd"], "end weight decay")
            self.wd_incr_steps = self._check_and_set(self.wd_incr_steps, sd["wd_incr_steps"],
                                                     "total number of weight decay iterations")
            self.wd_incr_style = self._check_and_set(self.wd_incr_style, sd["wd_incr_style"],
                                                     "weight decay incr style")
