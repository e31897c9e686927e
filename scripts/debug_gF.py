import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from gpu_common import *
from helpers import relerr
np.set_printoptions(precision=7, linewidth=200)
name = 'GatherMove-v1'
for trial in range(3):
    steps = 4
    scene, eng, o, o64 = make_pair(name, n=1200, substeps=1, max_steps=steps, twin=True)
    acts = actions_for(scene, steps, scale=1.0 / 19)
    n = eng.n_particles(); rng = np.random.RandomState(7)
    for s in range(steps):
        eng.set_action(s, acts[s][None]); eng.substep(s)
        for oo in (o, o64):
            oo.set_action(s, acts[s], n_substeps=1); oo.substep(s); sync_oracle_to_engine(eng, oo, s + 1, s + 1)
    for s in range(steps - 1, -1, -1):
        eng.zero_grad()
        gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
        gF, gC = f32(rng.normal(size=(n, 3, 3)) * 0.1), f32(rng.normal(size=(n, 3, 3)) * 1e-3)
        gt = f32(rng.normal(size=(eng.K, 8)) * 0.1)
        for i, t in enumerate(scene.tools):
            if t.state_dim == 7: gt[i, 7] = 0
        eng.add_particle_grad(s + 1, gx[None], gv[None], gF[None], gC[None]); eng.add_tool_grad(s + 1, gt[None])
        eng.substep_grad(s)
        a = eng.get_particle_grad(s)
        res = []
        for oo in (o, o64):
            oo.zero_grad(); oo.add_frame_grad(s + 1, gx, gv, gF, gC)
            for i in range(eng.K): oo.add_tool_grad(s + 1, i, gt[i])
            oo.substep_grad(s); res.append(oo.get_frame_grad(s))
        e = np.abs(a[2] - res[0][2]).reshape(n, -1).max(1)
        p = int(np.argmax(e))
        print(f'trial {trial} step {s}: gF err {relerr(a[2], res[0][2]):.2e} worst particle {p} abs {e[p]:.3e} scale {np.abs(res[0][2]).max():.3e}')
        if relerr(a[2], res[0][2]) > 1e-3:
            x, v, F, C = eng.get_particles(s)
            Ft, U, sg, V = o.get_svd()
            print('  F', F[p].ravel()); print('  C', C[p].ravel())
            print('  oracle32 Ftmp', Ft[p].ravel(), 'sig', sg[p])
            Ft64, U64, sg64, V64 = o64.get_svd()
            print('  oracle64 sig', sg64[p])
            Uc, sc, Vc = eng.debug_svd(f32(Ft[p])[None])
            print('  cuda sig', sc[0], 'U diff', np.abs(Uc[0] - U[p]).max(), 'V diff', np.abs(Vc[0] - V[p]).max())
            print('  gF cuda', a[2][p].ravel()); print('  gF o32 ', res[0][2][p].ravel()); print('  gF o64 ', res[1][2][p].ravel())
            yc = np.log(np.maximum(sg64[p], 0.05)); eh = yc - yc.mean()
            print('  delta_gamma64', np.sqrt((eh**2).sum() + 1e-8) - scene.yield_stress / (2 * scene.mu))
