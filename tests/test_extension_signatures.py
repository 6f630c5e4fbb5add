"""Python <-> C++ signature check that runs without a GPU: the built extension is imported on the CPU and every binding
whose argument list is assembled in Python (attention, fused TP, DP reduce, helpers) is called with CPU tensors of the
right types.  pybind11 validates the arguments before the function body runs, so a drift between a call site and its
binding shows up as TypeError here instead of on the GPU box; with matching arguments the call fails later (CUDA guard
on a CPU tensor / no driver), which is the expected outcome."""
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "megatron_llm_b200", "_C_b200.so")


@pytest.fixture(scope="module")
def ext():
    if not os.path.exists(SO):
        pytest.skip("extension not built (python -m megatron_llm_b200.ops.build)")
    spec = importlib.util.spec_from_file_location("_C_b200", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _accepts(fn, *args):
    try:
        fn(*args)
    except TypeError as e:        # pybind11 signature mismatch
        pytest.fail(f"{fn.__name__}: {str(e)[:600]}")
    except Exception:             # noqa: BLE001 - the body ran and tripped over the CPU tensors: arguments were accepted
        pass


def test_bindings_accept_the_python_call_sites(ext):
    bf = torch.bfloat16
    a = torch.zeros(256, 64, dtype=bf)
    w = torch.zeros(64, 64, dtype=bf)
    out = torch.zeros(256, 64, dtype=bf)
    i32 = torch.zeros(64, dtype=torch.int32)
    ptrs = [0, 0]
    # parallel/symm.py
    _accepts(ext.fused_ag_gemm, a, w, out, False, ptrs, 128, i32, i32, 0, ptrs, 0, 2, 1, 8, 0, 0, 0)
    _accepts(ext.fused_gemm_rs, a, w, out[:128], False, ptrs, 0, 128, 0, 4, i32, 0, ptrs, 0, 2, 1, 0, 0, [], 0)
    _accepts(ext.fused_gemm_rs, a, w, out[:128], False, ptrs, 0, 128, 0, 4, i32, 0, ptrs, 0, 2, 1, 0, 0, ptrs, 0)
    _accepts(ext.fused_ag_gemm_nvls, a, a[:128], w, out, False, 0, ptrs, i32, 128, 0, ptrs, 0, 2, 1, 8, 0, 0, 0)
    _accepts(ext.comm_copy2, a, a.clone(), a.clone())
    _accepts(ext.comm_set_state, i32, 0, 0, 0)
    _accepts(ext.dp_reduce, torch.zeros(64), ptrs, 0, ptrs, 0, 2, 1, 0.5, False, 8)
    _accepts(ext.p2p_bench, 0, 0, 0, 0, 1, 0, 0, 0, 0)
    # ops/attention_sm100.py
    q = torch.zeros(1, 128, 1, 128, dtype=bf)
    lse = torch.zeros(1, 1, 128)
    _accepts(ext.attn_fwd, q, q, q, True, -1, 0.088)
    _accepts(ext.attn_bwd, q, q, q, q, q, lse, True, -1, 0.088)
    mixed = torch.zeros(128, 1, 3 * 128, dtype=bf)
    ctx = torch.zeros(128, 1, 128, dtype=bf)
    _accepts(ext.attn_fwd_packed, mixed, 1, 1, -1, 0.088, 128)
    _accepts(ext.attn_bwd_packed, ctx, mixed, ctx, lse, 1, 1, -1, 0.088, 128)
    # ... with the dropout arguments (probability, 63-bit seed from ops._dropout_seed), fp16 tensors, the decode step
    seed = 0x7FFF_FFFF_FFFF_FFFF
    _accepts(ext.attn_fwd, q, q, q, True, -1, 0.088, 0.1, seed)
    _accepts(ext.attn_bwd, q, q, q, q, q, lse, True, -1, 0.088, 0.1, seed)
    _accepts(ext.attn_fwd_packed, mixed.half(), 1, 1, -1, 0.088, 128, 0.1, seed)
    _accepts(ext.attn_bwd_packed, ctx, mixed, ctx, lse, 1, 1, -1, 0.088, 128, 0.1, seed)
    _accepts(ext.attn_decode, q[:, :1], q, q, -1, 0.088, 0)
