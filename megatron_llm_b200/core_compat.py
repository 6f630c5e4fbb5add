"""``megatron.core`` namespace of the reference (core/__init__.py:1-12): ``parallel_state`` (alias ``mpu``),
``tensor_parallel`` and ``utils``."""
from . import parallel as tensor_parallel  # noqa: F401
from .parallel import state as parallel_state  # noqa: F401
from .parallel import state as mpu  # noqa: F401
from .utils import core_utils as utils  # noqa: F401

__all__ = ["parallel_state", "mpu", "tensor_parallel", "utils"]
