"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from diffskill_b200 import build as B
from diffskill_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    B.build()
    return engine.load_library()


def _declared():
    src = open(os.path.join(ROOT, 'include', 'diffskill_mpm.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dsk_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(engine.SYMBOLS) == names, set(engine.SYMBOLS) ^ set(names)


def test_abi_version_and_struct_layout(lib):
    assert lib.dsk_abi_version() == engine.ABI_VERSION
    assert lib.dsk_sizeof_config() == C.sizeof(engine.Config)
    assert lib.dsk_sizeof_tool_desc() == C.sizeof(engine.ToolDesc)
    assert lib.dsk_kernel_class_count() >= 10
    assert lib.dsk_kernel_class_name(2) == b'p2g'


def test_engine_fails_loudly_without_gpu(lib):
    """No CPU fallback: without a CUDA device the product path raises (on the GPU box this test is skipped)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from diffskill_b200.scene import load_scene
    scene, _ = load_scene('CutRearrange-v1')
    with pytest.raises(engine.EngineError):
        engine.Engine(scene, capacity=64, max_steps=1)


def test_sass_is_sm100a_with_vector_reductions():
    """The shipped cubin targets sm_100a and the scatter uses the 128-bit vector reduction."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    B.build()
    out = subprocess.run([cuobjdump, '-sass', engine.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    assert 'RED.E.ADD.F32x4' in out.replace('REDG', 'RED')
    assert 'SHFL.BFLY' in out
