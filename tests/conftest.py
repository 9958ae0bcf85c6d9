import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """SAME_B200_GUARD=1 python -m pytest tests -m gpu: the whole suite on canary-guarded, poisoned device buffers (the library's own
    memory checker, include/same_b200.h same_debug_guard).  Report the counters and fail the session if any zone was overwritten."""
    if os.environ.get("SAME_B200_GUARD", "0")[:1] != "1":
        return
    try:
        import torch
        if not torch.cuda.is_available():
            return
        from same_b200 import _lib as L
        bad, checked = L.debug_guard()
    except Exception as e:   # the library never loaded in this session
        print(f"\n[guard] no counters: {e}")
        return
    expected = int(os.environ.get("SAME_B200_GUARD_EXPECTED", "0"))   # deliberate overruns of the checker's own self-test
    print(f"\n[guard] device buffers checked on release: {checked}; canary zones found overwritten: {bad} "
          f"(of which {expected} on purpose, tests/test_gpu_guard.py)")
    if bad != expected:
        session.exitstatus = 1
