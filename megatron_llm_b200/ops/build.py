"""Ahead-of-time build of the in-tree sm_100a extension.

``python -m megatron_llm_b200.ops.build`` (or ``__graft_entry__.build()``) compiles every ``csrc/*.cu`` with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and links ``megatron_llm_b200/_C_b200.so``.  nvcc
cross-compiles without a GPU; the resulting .so travels with the source tree (no JIT at start-up, unlike
the reference's fused_kernels/__init__.py:17-98 which JIT-builds compute_70/80/90 only).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def build_main_extension(verbose: bool = False):
    from torch.utils import cpp_extension
    os.makedirs(BUILD, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*_bind.cpp")) +
                     [os.path.join(CSRC, "bindings.cpp")])
    sources = [s for s in sources if not os.path.basename(s).startswith("helpers")]
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 8))
    mod = cpp_extension.load(
        name="_C_b200", sources=sources, build_directory=BUILD,
        extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=NVCC_FLAGS,
        extra_ldflags=["-lcuda"], verbose=verbose, with_cuda=True)
    dst = os.path.join(PKG, "_C_b200.so")
    src = os.path.join(BUILD, "_C_b200.so")
    if (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst):
        # new inode, atomically: a process that has the old library mapped keeps running on the old file (copying over
        # it in place changes the pages under that process and crashes it -- seen with a test run during a rebuild)
        shutil.copy2(src, dst + ".tmp")
        os.replace(dst + ".tmp", dst)
    return mod


def build_helpers(verbose: bool = False):
    """CPU-only dataset index builders (pybind11, no torch / CUDA dependency)."""
    import sysconfig
    import pybind11
    src = os.path.join(CSRC, "helpers.cpp")
    if not os.path.exists(src):
        return None
    dst = os.path.join(PKG, "data", "_helpers_b200.so")
    if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
        return dst
    cmd = ["g++", "-O3", "-Wall", "-shared", "-std=c++17", "-fPIC", "-fdiagnostics-color",
           f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", src, "-o", dst]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return dst


def build_all(verbose: bool = False):
    build_helpers(verbose)
    return build_main_extension(verbose)


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv)
    print("built", os.path.join(PKG, "_C_b200.so"))
