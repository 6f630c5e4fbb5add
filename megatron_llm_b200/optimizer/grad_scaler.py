"""Loss scalers (parity: megatron/optimizer/grad_scaler.py: constant :38-50, dynamic with hysteresis :53-120)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from ..utils.device import current_device


class MegatronGradScaler(ABC):
    def __init__(self, initial_scale):
        assert initial_scale > 0.0
        self._scale = torch.tensor([initial_scale], dtype=torch.float, device=current_device())

    @property
    def scale(self):
        return self._scale

    @property
    def inv_scale(self):
        return self._scale.double().reciprocal().float()

    @abstractmethod
    def update(self, found_inf): ...

    @abstractmethod
    def state_dict(self): ...

    @abstractmethod
    def load_state_dict(self, state_dict): ...


class ConstantGradScaler(MegatronGradScaler):
    def update(self, found_inf):
        pass

    def state_dict(self):
        return dict()

    def load_state_dict(self, state_dict):
        pass


class DynamicGradScaler(MegatronGradScaler):
    """Halve (x backoff_factor) after ``hysteresis`` overflowing steps; grow after ``growth_interval`` clean ones."""

    def __init__(self, initial_scale, min_scale, growth_factor, backoff_factor, growth_interval, hysteresis):
        super().__init__(initial_scale)
        assert min_scale > 0.0 and min_scale <= initial_scale
        self.min_scale = torch.tensor([min_scale], dtype=torch.float, device=current_device())
        assert growth_factor > 1.0
        self.growth_factor = torch.tensor([growth_factor], dtype=torch.float, device=current_device())
        assert 0.0 < backoff_factor < 1.0
        self.backoff_factor = torch.tensor([backoff_factor], dtype=torch.float, device=current_device())
        assert growth_interval > 0
        self.growth_interval = growth_interval
        assert hysteresis > 0
        self.hysteresis = hysteresis
        self._growth_tracker = 0
        self._hysteresis_tracker = self.hysteresis

    def update(self, found_inf):
        if found_inf:
            self._growth_tracker = 0
            self._hysteresis_tracker -= 1
            if self._hysteresis_tracker <= 0:
                self._scale = torch.max(self._scale * self.backoff_factor, self.min_scale)
        else:
            self._growth_tracker += 1
            if self._growth_tracker == self.growth_interval:
                self._growth_tracker = 0
                self._hysteresis_tracker = self.hysteresis
                self._scale = self._scale * self.growth_factor

    def state_dict(self):
        return {"scale": self._scale, "growth_tracker": self._growth_tracker,
                "hysteresis_tracker": self._hysteresis_tracker}

    def load_state_dict(self, state_dict):
        self._scale = state_dict["scale"].to(current_device())
        self._growth_tracker = state_dict["growth_tracker"]
        self._hysteresis_tracker = state_dict["hysteresis_tracker"]
