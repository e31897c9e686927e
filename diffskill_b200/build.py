"""Builds libdiffskill_mpm.so (the C-ABI CUDA library) in-tree for sm_100a.

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'libdiffskill_mpm.so')
SO_TIMELINE = os.path.join(HERE, 'libdiffskill_mpm_tl.so')   # profiling build, see dsk_timeline_* in the header
SO_PRECISE = os.path.join(HERE, 'libdiffskill_mpm_pm.so')    # diagnostic build: correctly rounded log/exp/div/rsqrt (DSK_PRECISE_MATH)
SOURCES = ['engine.cu']
HEADERS = ['mpm_math.cuh', 'svd3.cuh', 'tools.cuh', 'particle_math.cuh', 'kernels_common.cuh', 'kernels_aux.cuh', 'kernels_fwd.cuh',
           'kernels_bwd.cuh', os.path.join('..', '..', 'include', 'diffskill_mpm.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


def _host_cxx():
    return '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else (shutil.which('g++') or 'g++')


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, extra=(), timeline=False, precise=False):
    """timeline=True builds the profiling variant (kernels stamp %globaltimer) next to the product library;
    precise=True the diagnostic variant without fast-math transcendentals."""
    if precise:
        cmd = [_nvcc(), '-ccbin', _host_cxx()] + NVCC_FLAGS + ['-DDSK_PRECISE_MATH'] + list(extra) + ['-o', SO_PRECISE] + \
            [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
        return SO_PRECISE
    if timeline:
        cmd = [_nvcc(), '-ccbin', _host_cxx()] + NVCC_FLAGS + ['-DDSK_TIMELINE'] + list(extra) + ['-o', SO_TIMELINE] + \
            [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
        return SO_TIMELINE
    if not force and not needs_build():
        return SO
    cmd = [_nvcc(), '-ccbin', _host_cxx()] + NVCC_FLAGS + list(extra) + ['-o', SO] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True, extra=['-Xptxas', '-v'] if '--ptxas' in sys.argv else [],
          timeline='--timeline' in sys.argv, precise='--precise' in sys.argv)
