import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if not os.path.exists("/dev/nvidiactl"):
    # CPU box: most tests start 2-4 ranks on matrices of a few KB, each of which would otherwise spin up one OpenMP
    # thread per core; two threads per process makes the suite ~15 % faster (16.5 -> 15 min) and the timing of the
    # multi-process tests less sensitive to the core count.  (Inherited by the ranks the tests launch.)
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    os.environ.setdefault("MKL_NUM_THREADS", "2")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "multigpu: test needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
        ngpu = torch.cuda.device_count() if has_cuda else 0
    except Exception:
        has_cuda, ngpu = False, 0
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_multi = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    for item in items:
        if "gpu" in item.keywords and not has_cuda:
            item.add_marker(skip_gpu)
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(skip_multi)
