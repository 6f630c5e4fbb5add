"""Loader for the in-tree sm_100a extension (``megatron_llm_b200/_C_b200.so``).

The extension is built ahead of time by ``__graft_entry__.build()`` /
``python -m megatron_llm_b200.ops.build`` with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and travels with the source tree.  There
is no JIT at start-up (the reference JIT-compiles its fused kernels on rank 0 behind a barrier,
fused_kernels/__init__.py:17-98) and no silent fallback: on a CUDA device a missing extension
is a hard error.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import threading

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT_NAME = "_C_b200"
_lock = threading.Lock()
_mod = None
_err = None

# number of hand-written kernels launched (bench.py reports this as ``gpu_launches``)
LAUNCHES = 0


def count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def so_path() -> str:
    return os.path.join(_PKG_DIR, EXT_NAME + ".so")


def load():
    """Import the prebuilt extension; raises with a clear message if missing."""
    global _mod, _err
    if _mod is not None:
        return _mod
    with _lock:
        if _mod is not None:
            return _mod
        path = so_path()
        if not os.path.exists(path):
            _err = (f"sm_100a extension not found at {path}; run `python -c \"import __graft_entry__ as g; "
                    f"g.build()\"` (or python -m megatron_llm_b200.ops.build) first")
            raise RuntimeError(_err)
        loader = importlib.machinery.ExtensionFileLoader(EXT_NAME, path)
        spec = importlib.util.spec_from_file_location(EXT_NAME, path, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _mod = mod
        return _mod


def available() -> bool:
    try:
        load()
        return True
    except Exception:
        return False


def cuda_ops_available(t: torch.Tensor) -> bool:
    """True -> run the hand-written kernel.  CPU tensors use the torch oracle; a CUDA tensor
    with no extension is an error (never a silent library fallback)."""
    if not t.is_cuda:
        return False
    if os.environ.get("MLB200_DISABLE_KERNELS", "0") == "1":
        return False
    load()
    return True
