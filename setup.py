"""Packaging (parity: setup.py of the reference, which ships only ``megatron.core``; this one ships the whole package).

``python setup.py build_ext --inplace`` (or ``pip install -e .``) compiles the sm_100a extension next to the sources
(``megatron_llm_b200/_C_b200.so``) with the same flags as ``python -m megatron_llm_b200.ops.build`` and the C++ dataset
helpers."""
import os
import subprocess
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_ext import build_ext as _build_ext

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildNative(_build_ext):
    """Delegates to the in-tree builder (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo; no GPU needed)."""

    def run(self):
        subprocess.check_call([sys.executable, "-m", "megatron_llm_b200.ops.build"], cwd=ROOT)


class BuildHelpers(Command):
    description = "build the C++ dataset index helpers only"
    user_options = []

    def initialize_options(self):
        pass

    def finalize_options(self):
        pass

    def run(self):
        sys.path.insert(0, ROOT)
        from megatron_llm_b200.ops.build import build_helpers
        build_helpers()


def readme():
    with open(os.path.join(ROOT, "README.md"), encoding="utf-8") as f:
        return f.read()


setup(
    name="megatron_llm_b200",
    version="0.1.0",
    description="Blackwell-native 3D-parallel LLM training framework with the capabilities of epfLLM/Megatron-LLM",
    long_description=readme(),
    long_description_content_type="text/markdown",
    python_requires=">=3.10",
    packages=find_packages(include=("megatron_llm_b200", "megatron_llm_b200.*", "megatron", "megatron.*",
                                    "weights_conversion", "weights_conversion.*", "tasks", "tasks.*")),
    package_data={"megatron_llm_b200": ["csrc/*", "static/*", "*.so"]},
    install_requires=["torch>=2.4", "numpy", "regex", "sentencepiece"],
    extras_require={"hf": ["transformers"], "logging": ["wandb", "tensorboard"]},
    cmdclass={"build_ext": BuildNative, "build_helpers": BuildHelpers},
    classifiers=["Programming Language :: Python :: 3", "Environment :: GPU :: NVIDIA CUDA :: 12",
                 "Topic :: Scientific/Engineering :: Artificial Intelligence"],
)
