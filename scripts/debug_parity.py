"""Diagnostic: where do CUDA and oracle differ after one substep? (run on the GPU box)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from gpu_common import *
from helpers import relerr

for name in ENVS:
    scene, eng, o = make_pair(name, n=1500, substeps=1, max_steps=2)
    acts = actions_for(scene, 1, scale=1.0 / 19)
    eng.set_action(0, acts[0][None]); o.set_action(0, acts[0], n_substeps=1)
    eng.substep(0); o.substep(0)
    _, vout, m, _ = eng.debug_grid()
    ovin, ovout, om = o.get_grid()
    n = scene.n_grid
    err = np.abs(vout - ovout).max(-1)
    scale = np.abs(ovout).max()
    occ = om > 1e-12
    # distance to nearest tool (oracle probe, frame 0)
    idx = np.argwhere(occ)
    sd = np.full((n, n, n), 9.0)
    for I in idx:
        p = I / n
        sd[tuple(I)] = min(o.tool_sdf(t, 0, p.astype(np.float32).astype(np.float64)) for t in range(eng.K))
    far = occ & (sd > 0.02)
    near = occ & (sd <= 0.02)
    print(f'== {name}: scale {scale:.3f}  nodes {occ.sum()}  far {far.sum()} near {near.sum()}')
    print('   rel err far  %.2e   near %.2e   m err %.2e' % (err[far].max() / scale if far.any() else 0, err[near].max() / scale if near.any() else 0, relerr(m, om)))
    order = np.argsort(-err.ravel())[:6]
    for fl in order:
        I = np.unravel_index(fl, err.shape)
        p = np.array(I) / n
        sds = [o.tool_sdf(t, 0, p) for t in range(eng.K)]
        print('   node', I, 'm %.3e' % om[I], 'err %.2e' % err[I], 'v_cuda', vout[I], 'v_orc', ovout[I], 'v0', ovin[I] / om[I], 'sdf', ['%.4f' % s for s in sds])
