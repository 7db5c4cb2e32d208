import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    if os.environ.get("KRY_TEST_DOUBLE"):
        # Development aid for a container without a GPU: drive the SOLVER-level GPU test modules over
        # the numpy test double of the device layer (tests/fake_device.py) to catch host-logic
        # regressions and stale expectations before GPU time is spent, e.g.
        #   KRY_TEST_DOUBLE=1 python -m pytest tests/test_solvers_gpu.py -m gpu -q
        # Kernel-level tests need the real device (their ctx fixture asserts CUDA); the tight
        # tolerances of the kappa=1e5 cases assume the device's arithmetic.  Proves nothing about kernels.
        import fake_device
        from krypy_b200 import _device
        fake = fake_device.FakeContext()
        _device.Context.get = classmethod(lambda cls, device=None: fake)
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
