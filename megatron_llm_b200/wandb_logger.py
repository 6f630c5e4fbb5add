"""Weights & Biases logging behind a TensorBoard-like shim (parity: megatron/wandb_logger.py:14-174).

``WandbTBShim`` exposes ``add_scalar / add_text / flush_all``; scalars of one step are buffered and committed
together when the step changes; the "vs samples" duplicate series are dropped.  If ``wandb`` cannot be
imported or initialised (no network) it degrades to the TensorBoard mirror only."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Dict, Optional


@dataclass
class WandBConfig:
    project: str = "megatron-llm-b200"
    entity: Optional[str] = None
    name: Optional[str] = None
    id: Optional[str] = None
    resume: str = "allow"
    api_key: Optional[str] = None
    config_args: Dict[str, Any] = field(default_factory=dict)
    with_tensorboard: bool = True
    logdir: Optional[str] = None
    try_catch_guard: bool = True

    @staticmethod
    def default(project, run_id=None):
        return WandBConfig(project=project, id=run_id)

    @staticmethod
    def from_args(args) -> "WandBConfig":
        assert args.wandb_logger
        return WandBConfig(project=args.wandb_project or "megatron-llm-b200", entity=args.wandb_entity,
                           name=args.wandb_name, id=args.wandb_id, resume=args.wandb_resume,
                           api_key=args.wandb_api_key,
                           config_args={k: str(v) for k, v in vars(args).items()},
                           with_tensorboard=bool(args.tensorboard_dir), logdir=args.tensorboard_dir)


def _import_wandb():
    try:
        import wandb
        return wandb
    except Exception:
        return None


class WandbTBShim:
    """Drop-in for ``SummaryWriter`` that forwards to wandb (and optionally mirrors to TensorBoard)."""

    def __init__(self, config: WandBConfig):
        self.cfg = config
        self._wandb = _import_wandb()
        self._run = None
        if config.api_key:
            os.environ.setdefault("WANDB_API_KEY", config.api_key)
        if self._wandb is not None:
            try:
                self._run = self._wandb.init(project=config.project, entity=config.entity, name=config.name,
                                             id=config.id, resume=config.resume, config=config.config_args,
                                             mode=os.environ.get("WANDB_MODE", "offline"))
            except Exception as e:
                if not config.try_catch_guard:
                    raise
                print(f"WARNING: wandb.init failed ({e!r}); logging to TensorBoard only", flush=True)
                self._run = None
        self._tb = None
        if config.with_tensorboard and config.logdir:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._tb = SummaryWriter(log_dir=config.logdir)
            except Exception:
                self._tb = None
        self._step: Optional[int] = None
        self._pending: Dict[str, Any] = {}

    def _commit(self):
        if self._pending and self._run is not None:
            try:
                self._wandb.log(self._pending, step=self._step, commit=True)
            except Exception as e:
                if not self.cfg.try_catch_guard:
                    raise
                print(f"WARNING: wandb.log failed ({e!r})", flush=True)
        self._pending = {}

    def add_scalar(self, name: str, var, step: int):
        if self._tb is not None:
            self._tb.add_scalar(name, var, global_step=step)
        if " vs " in name:   # "x vs samples" duplicates are TensorBoard-only
            return
        if self._step is not None and step != self._step:
            self._commit()
        self._step = step
        self._pending[name] = float(var) if hasattr(var, "__float__") else var

    def add_text(self, name: str, value: str, global_step: int = 0):
        if self._tb is not None:
            self._tb.add_text(name, value, global_step=global_step)
        if self._run is not None:
            try:
                self._run.config.update({name: value}, allow_val_change=True)
            except Exception:
                pass

    def flush_all(self):
        self._commit()
        if self._tb is not None:
            self._tb.flush()

    flush = flush_all
