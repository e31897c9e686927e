"""Diagnostic (GPU box): one substep forward + adjoint of a scene variant, CUDA vs fp32 / fp64 oracle per quantity."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from gpu_common import *
from helpers import relerr, small_dough, perturbed_state, tool_start
from diffskill_b200.engine import Engine
from oracle import oracle as orc
np.set_printoptions(precision=5, linewidth=200)


def run(name, keep=None, lift=None, n=1200, steps=2, label=''):
    scene, cfg, x0 = small_dough(name, n)
    scene = copy.deepcopy(scene); scene.substeps = 1
    st0 = [f32(s) for s in tool_start(name, scene)]
    if lift is not None:
        st0[lift[0]][1] = lift[1]
    if keep is not None:
        scene.tools = [scene.tools[i] for i in keep]; st0 = [st0[i] for i in keep]
    v0, F0, C0 = perturbed_state(x0, 1)
    x0, v0, F0, C0 = f32(x0), f32(v0), f32(F0), f32(C0)
    eng = Engine(scene, n_envs=1, capacity=n, max_steps=steps)
    eng.set_particles(0, 0, x0, v0, F0, C0)
    os_ = [orc.Oracle(scene, n, steps + 1, f64=f, threads=1) for f in (False, True)]
    for o in os_:
        o.set_frame(0, x0, v0, F0, C0)
    for i, s in enumerate(st0):
        eng.set_tool_state(0, 0, i, s)
        for o in os_: o.set_tool_state(0, i, s)
    acts = actions_for(scene, steps, scale=1.0 / 19) if scene.action_dim else np.zeros((steps, 0), np.float32)
    for s in range(steps):
        if scene.action_dim: eng.set_action(s, acts[s][None])
        eng.substep(s)
        for o in os_:
            if scene.action_dim: o.set_action(s, acts[s], n_substeps=1)
            o.substep(s)
        x, v, F, C = eng.get_particles(s + 1)
        fe = [[relerr(a, b) for a, b in zip((x, v, F, C), o.get_frame(s + 1))] for o in os_]
        print(f'  [{label}] fwd substep {s}: vs f32 x %.1e v %.1e F %.1e C %.1e | vs f64 x %.1e v %.1e F %.1e C %.1e' % tuple(fe[0] + fe[1]))
        for o in os_: sync_oracle_to_engine(eng, o, s + 1, s + 1)
    rng = np.random.RandomState(7)
    for s in range(steps - 1, -1, -1):
        gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
        gF, gC = f32(rng.normal(size=(n, 3, 3)) * 0.1), f32(rng.normal(size=(n, 3, 3)) * 1e-3)
        eng.zero_grad()
        eng.add_particle_grad(s + 1, gx[None], gv[None], gF[None], gC[None])
        eng.substep_grad(s)
        a = eng.get_particle_grad(s)
        ga_in, gm = eng.debug_grid_grad()
        mine = dict(gx=a[0], gv=a[1], gF=a[2], gC=a[3], grid_gv=ga_in, grid_gm=gm, tool=eng.get_tool_grads(s))
        if scene.action_dim: mine['action'] = eng.get_action_grad(s)[0]
        res = []
        for o in os_:
            o.zero_grad(); o.add_frame_grad(s + 1, gx, gv, gF, gC); o.substep_grad(s)
            if scene.action_dim: o.L.orc_set_velocity_grad(o.h, s, 1)
            b = o.get_frame_grad(s); og, _, ogm = o.get_grid_grad()
            d = dict(gx=b[0], gv=b[1], gF=b[2], gC=b[3], grid_gv=og, grid_gm=ogm, tool=o.get_tool_grads(s))
            if scene.action_dim: d['action'] = o.get_action_grad(s)
            res.append(d)
        print(f'  [{label}] bwd substep {s}:', {k: '%.1e/%.1e (floor %.1e)' % (relerr(mine[k], res[0][k]), relerr(mine[k], res[1][k]), relerr(res[0][k], res[1][k])) for k in mine})
        if s == steps - 1:
            print('    tool cuda', np.asarray(mine['tool']).ravel()); print('    tool o64 ', np.asarray(res[1]['tool']).ravel())
            gv_ = np.asarray(mine['grid_gv']).reshape(-1, 3); ov_ = np.asarray(res[1]['grid_gv']).reshape(-1, 3)
            bad = np.argsort(-np.abs(gv_ - ov_).max(1))[:4]
            ng = scene.n_grid
            for b_ in bad:
                print('    worst node', (b_ // (ng * ng), (b_ // ng) % ng, b_ % ng), 'cuda', gv_[b_], 'o64', ov_[b_])


def safe(*a, **k):
    try:
        run(*a, **k)
    except Exception as e:
        print('  FAILED', k.get('label'), repr(e))


if __name__ == '__main__':
    run('Rope-v1', label='rope full')
    run('Rope-v1', keep=[2], label='rope cylinder only')
    run('Rope-v1', keep=[0, 1], label='rope spheres only')
    run('Rope-v1', keep=[2], lift=(2, 0.05), label='rope cylinder lifted to y=0.05')
    safe('Rope-v1', keep=[], label='rope no tools')
    run('Torus-v1', label='torus')
    safe('Torus-v1', keep=[], label='torus no tools')
