import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: minutes-long CPU test")


def _gpu_ready():
    """a CUDA device of compute capability 10.x and the built library (the product has no CPU fallback, so without
    them the `gpu` tests cannot do anything but fail in lr_device_check)"""
    try:
        import torch
        if not torch.cuda.is_available() or torch.cuda.get_device_capability(0)[0] != 10:
            return False, "no sm_100 CUDA device"
        from llava_reward_b200 import _lib
        if not os.path.exists(_lib.LIB_PATH):
            return False, f"{_lib.LIB_PATH} not built"
        return True, ""
    except Exception as e:  # pragma: no cover
        return False, repr(e)


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for it in gpu_items:
        it.add_marker(skip)
