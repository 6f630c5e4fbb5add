"""Loader for the C++ index builders (``csrc/helpers.cpp`` -> ``_helpers_b200.so``); builds on demand with g++
(the reference ships a Makefile hard-wired to python3.10-config, data/Makefile:4)."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_helpers_b200.so")
_mod = None


def _load():
    global _mod
    if _mod is not None:
        return _mod
    if not os.path.exists(_SO):
        from ..ops.build import build_helpers
        build_helpers()
    loader = importlib.machinery.ExtensionFileLoader("_helpers_b200", _SO)
    spec = importlib.util.spec_from_file_location("_helpers_b200", _SO, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mod = mod
    return mod


def build_sample_idx(*a):
    return _load().build_sample_idx(*a)


def build_blending_indices(*a):
    return _load().build_blending_indices(*a)


def build_mapping(*a):
    return _load().build_mapping(*a)


def build_blocks_mapping(*a):
    return _load().build_blocks_mapping(*a)
