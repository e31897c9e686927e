"""CPU study (no GPU): how far does a scene's multi-step gradient move under the freedoms the CUDA build takes relative to
the oracle -- hardware log / exp / div / rsqrt (the reference itself runs fast_math=True) and FMA contraction?

The CPU twin of the engine's substep and adjoint substep (tests/host_check/host_twin.cpp: the product's device functions
compiled by g++) is chained over a whole H-step rollout, forward and backward, in several builds: exact intrinsics without
and with FMA contraction, and approximate intrinsics (pseudo-random errors of the size the hardware is allowed: log 2^-21.4
absolute, exp / div / rsqrt 2 ulp), one family at a time and all together.  x.grad[0] of each is compared with the fp32 / fp64 oracle.  Tool poses come from the oracle (the kinematics
do not depend on the dough).

    python scripts/fastmath_sensitivity.py Rope-v1 LiftSpread-v1 ...
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import perturbed_state, relerr, small_dough, tool_start  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from diffskill_b200.engine import make_config  # noqa: E402

HC = os.path.join(ROOT, 'tests', 'host_check')
FP = C.POINTER(C.c_float)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(FP)


def build(tag, flags):
    so = os.path.join('/tmp', f'libhost_twin_{tag}.so')
    subprocess.check_call(['g++', '-std=c++17', '-fPIC', '-shared', '-w'] + flags + ['-o', so, os.path.join(HC, 'host_twin.cpp')])
    return C.CDLL(so)


def run(name, twins, H=3, n=800):
    scene, cfg, x0 = small_dough(name, n, 0)
    v0, F0, C0 = perturbed_state(x0, 1)
    S = scene.substeps
    st0 = [f32(s) for s in tool_start(name, scene)]
    acts = f32(np.random.RandomState(3).uniform(-1, 1, (H, scene.action_dim)) * 0.7)
    rng = np.random.RandomState(11)
    gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
    refs, oracles = [], []
    for f64 in (False, True):
        o = orc.Oracle(scene, n, H * S + 2, f64=f64, threads=8)
        o.set_frame(0, f32(x0), f32(v0), f32(F0), f32(C0))
        for i, s in enumerate(st0):
            o.set_tool_state(0, i, s)
        for s in range(H):
            o.forward_step(s, acts[s])
        o.zero_grad()
        o.add_frame_grad(H * S, gx, gv)
        for s in range(H - 1, -1, -1):
            o.backward_step(s)
        refs.append((o.get_frame_grad(0)[0], o.get_frame(H * S)))
        oracles.append(o)
    o32 = oracles[0]
    cfgc = make_config(scene, 1, n, 1, 1, True, 666., 0)
    K, G = len(scene.tools), scene.n_grid ** 3
    mat = f32(np.stack([np.full(n, scene.mu), np.full(n, scene.lam), np.full(n, scene.yield_stress)]))
    poses = [f32(np.stack([[o32.get_tool_state(ff, i) for i in range(K)] for ff in (f, f + 1)])) for f in range(H * S)]
    sc = np.abs(refs[1][0]).max()
    d32 = np.abs(refs[0][0] - refs[1][0]).max(1)
    print(f'{name}: fp32 oracle vs fp64: x.grad[0] %.1e (%d of %d particles above 1e-4 of the max), v %.1e' %
          (relerr(refs[0][0], refs[1][0]), (d32 > 1e-4 * sc).sum(), n, relerr(refs[0][1][1], refs[1][1][1])))
    results = {}
    for tag, tw in twins:
        fr = [(f32(x0), f32(v0), f32(np.reshape(C0, (n, 9))), f32(np.reshape(F0, (n, 9))))]
        grids = []
        for f in range(H * S):
            x, v, Cm, F = fr[-1]
            xn, vn, Cn, Fn = (np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 9), np.float32),
                              np.zeros((n, 9), np.float32))
            g0, gvo = np.zeros((G, 4), np.float32), np.zeros((G, 3), np.float32)
            tw.hc_substep(C.byref(cfgc), n, _p(x), _p(v), _p(Cm), _p(F), _p(mat), _p(poses[f]), 1, _p(xn), _p(vn), _p(Cn),
                          _p(Fn), _p(g0), _p(gvo))
            fr.append((xn, vn, Cn, Fn))
            grids.append((g0, gvo))
        a = [gx.copy(), gv.copy(), np.zeros((n, 9), np.float32), np.zeros((n, 9), np.float32)]
        for f in range(H * S - 1, -1, -1):
            x, v, Cm, F = fr[f]
            o_ = [np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 9), np.float32),
                  np.zeros((n, 9), np.float32)]
            ga, padj = np.zeros((G, 4), np.float32), np.zeros((2, max(K, 1), 8), np.float32)
            tw.hc_substep_grad(C.byref(cfgc), n, _p(x), _p(v), _p(Cm), _p(F), _p(fr[f + 1][0]), _p(mat), _p(grids[f][0]),
                               _p(grids[f][1]), _p(poses[f]), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(o_[0]), _p(o_[1]),
                               _p(o_[2]), _p(o_[3]), _p(ga), _p(padj))
            a = o_
        d = np.abs(a[0] - refs[1][0]).max(1)
        print(f'  twin [{tag:34s}] x.grad[0] vs fp32 oracle %.1e, vs fp64 %.1e (%3d particles above 1e-4 of the max); '
              'state after %d substeps: v %.1e F %.1e vs fp32 oracle' %
              (relerr(a[0], refs[0][0]), relerr(a[0], refs[1][0]), (d > 1e-4 * sc).sum(), H * S,
               relerr(fr[-1][1], refs[0][1][1]), relerr(fr[-1][3].reshape(n, 3, 3), refs[0][1][2])))
        results[tag] = dict(gx_vs_f32=relerr(a[0], refs[0][0]), gx_vs_f64=relerr(a[0], refs[1][0]),
                            floor=relerr(refs[0][0], refs[1][0]), particles_off=int((d > 1e-4 * sc).sum()),
                            v_vs_f32=relerr(fr[-1][1], refs[0][1][1]), x_vs_f32=relerr(fr[-1][0], refs[0][1][0]))
    return results


if __name__ == '__main__':
    twins = [('exact intrinsics, no contraction', build('exact', ['-O1', '-ffp-contract=off'])),
             ('exact intrinsics, FMA contraction', build('fma', ['-O2', '-ffp-contract=fast', '-mfma'])),
             ('approximate log only', build('log', ['-O1', '-DHC_APPROX=1', '-ffp-contract=off'])),
             ('approximate exp only', build('exp', ['-O1', '-DHC_APPROX=2', '-ffp-contract=off'])),
             ('approximate div + rsqrt (SVD) only', build('svd', ['-O1', '-DHC_APPROX=4', '-ffp-contract=off'])),
             ('all approximate + FMA contraction', build('approx_fma', ['-O2', '-DHC_APPROX=7', '-ffp-contract=fast', '-mfma'])),
             ('SVD approx., BIASED cosine (round 1)', build('svd_b', ['-O1', '-DHC_APPROX=4', '-DDSK_BIASED_COSINE', '-ffp-contract=off'])),
             ('all approx. + FMA, BIASED cosine', build('approx_fma_b', ['-O2', '-DHC_APPROX=7', '-DDSK_BIASED_COSINE',
                                                                        '-ffp-contract=fast', '-mfma']))]
    for nm in (sys.argv[1:] or ['Rope-v1', 'LiftSpread-v1']):
        run(nm, twins)
