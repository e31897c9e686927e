"""TEST INFRASTRUCTURE: builds libdiffskill_mpm_emu.so, the engine of diffskill_b200/csrc compiled for the CPU.

engine.cu is copied with two mechanical edits -- `#include <cuda_runtime.h>` dropped (tests/host_check/cuda_rt_shim.h stands
in for it) and every `kernel<<<grid, block, smem, stream>>>(args)` rewritten to `simt_launch(grid, block, smem, [&] {
kernel(args); })` -- and compiled by g++ together with the unmodified kernel headers on top of the thread-block emulation of
simt_shim.h.  The result exports the same C ABI (include/diffskill_mpm.h): the `-m gpu` parity cases can be driven through
it on a machine without a GPU (DSK_LIB=emu, graphs off), slowly, which is what tests/test_emulated_engine.py does for a
few of them.  It shares every line of kernel and scheduling code with the product except the CUDA-graph capture path; its
arithmetic is the host's (correctly rounded sqrt / div / log / exp), not the GPU's.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'diffskill_b200', 'csrc')
GEN = os.path.join(HERE, '_gen')
SO = os.path.join(HERE, 'libdiffskill_mpm_emu.so')


def rewrite_launches(src):
    out, i = [], 0
    while True:
        m = src.find('<<<', i)
        if m < 0:
            out.append(src[i:])
            break
        # kernel name (identifier with optional template arguments) ends right before '<<<'
        j = m
        if src[j - 1] == '>':
            depth = 0
            while True:
                j -= 1
                if src[j] == '>':
                    depth += 1
                elif src[j] == '<':
                    depth -= 1
                    if depth == 0:
                        break
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] == '_'):
            j -= 1
        name = src[j:m]
        e = src.index('>>>', m)
        cfg = src[m + 3:e]
        assert src[e + 3] == '(', src[m - 40:e + 10]
        k, depth = e + 3, 0
        while True:
            if src[k] == '(':
                depth += 1
            elif src[k] == ')':
                depth -= 1
                if depth == 0:
                    break
            k += 1
        args = src[e + 4:k]
        parts = [p.strip() for p in re.split(r',(?![^()]*\))', cfg)]
        assert len(parts) in (2, 3, 4), cfg
        smem = parts[2] if len(parts) > 2 else '0'
        out.append(src[i:j])
        out.append(f'simt_launch(dim3({parts[0]}), dim3({parts[1]}), (size_t)({smem}), [&] {{ {name}({args}); }})')
        i = k + 1
    return ''.join(out)


def build(verbose=False, flags=(), so=None):
    """flags / so: study variants (e.g. -DHC_APPROX=7 -ffp-contract=fast -mfma: GPU-like arithmetic, see cuda_shim.h)."""
    so = so or SO
    os.makedirs(GEN, exist_ok=True)
    src = open(os.path.join(CSRC, 'engine.cu')).read()
    src = src.replace('#include <cuda_runtime.h>', '// (cuda_runtime.h: tests/host_check/cuda_rt_shim.h through mpm_math.cuh)')
    src = src.replace('#include "kernels_bwd.cuh"', f'#include "{os.path.join(CSRC, "kernels_bwd.cuh")}"')
    src = src.replace('#include "../../include/diffskill_mpm.h"', f'#include "{os.path.join(ROOT, "include", "diffskill_mpm.h")}"')
    # persistent grids are sized for 148 SMs; the emulation walks the blocks of a launch one after the other, so it uses a
    # handful (every kernel is a grid-stride loop: the grid size is a launch parameter, not part of the algorithm)
    for old, new in (('return e->big ? 148 * 8 : 148 * e->grid_ctas_per_sm;', 'return 3;'), ('k_grid_flat<<<148 * (e)->flat_fwd_ctas_per_sm,', 'k_grid_flat<<<3,'),
                     ('k_grid_adj_flat<<<148 * (e)->flat_adj_ctas_per_sm,', 'k_grid_adj_flat<<<3,')):
        assert src.count(old) == 1, old
        src = src.replace(old, new)
    gen = os.path.join(GEN, 'engine_emu.cpp')
    open(gen, 'w').write('#define DSK_HOST_CHECK 1\n#define DSK_HOST_SIMT 1\n#define DSK_HOST_EMU 1\n' + rewrite_launches(src))
    cmd = ['g++', '-O1', '-std=c++17', '-fPIC', '-shared', '-pthread', '-w', '-U_FORTIFY_SOURCE', '-D_FORTIFY_SOURCE=0'] + \
        (list(flags) or ['-ffp-contract=off']) + ['-o', so, gen]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return so


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.h', '.py'))]
    return any(os.path.getmtime(d) > t for d in deps)


if __name__ == '__main__':
    build(verbose=True)
