import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    if os.environ.get('DSK_LIB') == 'emu':
        # TEST INFRASTRUCTURE ONLY: point the binding at the CPU emulation of the engine (tests/host_check/make_emu.py).  The
        # product loader (diffskill_b200/engine.py) knows nothing about it -- the switch lives here, in the test harness.
        os.environ['DSK_NO_GRAPHS'] = '1'
        from diffskill_b200 import engine
        engine.LIB_PATH = os.environ.get('DSK_EMU_LIB') or os.path.join(ROOT, 'tests', 'host_check', 'libdiffskill_mpm_emu.so')


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
