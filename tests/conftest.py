import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    if os.environ.get('DSK_LIB') == 'emu':     # CPU emulation of the engine (tests/host_check/make_emu.py) stands in for it
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if any(i.get_closest_marker('gpu') for i in items) and not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device in this container; GPU tests run under gpurun")
        for i in items:
            if i.get_closest_marker('gpu'):
                i.add_marker(skip)
