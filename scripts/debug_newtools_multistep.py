"""Diagnostic (GPU box): H env steps of the real substep count, action gradient and x.grad[0] of scene variants,
CUDA vs fp32 / fp64 oracle (is a multi-step mismatch a kernel bug or the fp32 noise floor of the scene?)."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from gpu_common import *
from helpers import relerr, small_dough, perturbed_state, tool_start
from diffskill_b200.engine import Engine
from oracle import oracle as orc
np.set_printoptions(precision=4, linewidth=220)


def run(name, keep=None, H=3, n=800, label='', ys=None, gf=None, slots=1, sort=True, act_scale=0.7, lift=None, S_=None,
        verbose=True):
    scene, cfg, x0 = small_dough(name, n)
    scene = copy.deepcopy(scene)
    if ys is not None: scene.yield_stress = ys
    if gf is not None: scene.ground_friction = gf
    if S_ is not None: scene.substeps = S_
    st0 = [f32(s) for s in tool_start(name, scene)]
    for i_, y_ in (lift or []):
        st0[i_][1] = y_
    if keep is not None:
        scene.tools = [scene.tools[i] for i in keep]; st0 = [st0[i] for i in keep]
    v0, F0, C0 = perturbed_state(x0, 1)
    x0, v0, F0, C0 = f32(x0), f32(v0), f32(F0), f32(C0)
    S = scene.substeps
    eng = Engine(scene, n_envs=1, capacity=n, max_steps=H, step_slots=slots, sort=sort)
    eng.set_particles(0, 0, x0, v0, F0, C0)
    os_ = [orc.Oracle(scene, n, H * S + 1, f64=f, threads=8) for f in (False, True)]
    for o in os_: o.set_frame(0, x0, v0, F0, C0)
    for i, s in enumerate(st0):
        eng.set_tool_state(0, 0, i, s)
        for o in os_: o.set_tool_state(0, i, s)
    A = scene.action_dim
    acts = actions_for(scene, H, scale=act_scale) if A else np.zeros((H, 0), np.float32)
    for s in range(H):
        if A: eng.set_action(s, acts[s][None])
        eng.forward_step(s)
        for o in os_:
            if A: o.forward_step(s, acts[s])
            else:
                for j in range(s * S, (s + 1) * S): o.substep(j)
    x, v, F, C = eng.get_particles(H)
    for o, nm in zip(os_, ('f32', 'f64')):
        ox, ov, oF, oC = o.get_frame(H * S)
        print(f'  [{label}] state vs {nm}: x %.1e v %.1e F %.1e C %.1e' % (relerr(x, ox), relerr(v, ov), relerr(F, oF), relerr(C, oC)))
    rng = np.random.RandomState(11)
    gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
    eng.zero_grad(); eng.add_particle_grad(H, gx[None], gv[None])
    ga = np.zeros((H, A))
    for s in range(H - 1, -1, -1):
        eng.backward_step(s)
        if A: ga[s] = eng.get_action_grad(s)[0]
    a = eng.get_particle_grad(0)[0]
    ogs, oxs = [], []
    for o in os_:
        o.zero_grad(); o.add_frame_grad(H * S, gx, gv)
        og = np.zeros((H, A))
        for s in range(H - 1, -1, -1):
            if A: og[s] = o.backward_step(s)
            else:
                for j in range((s + 1) * S - 1, s * S - 1, -1): o.substep_grad(j)
        ogs.append(og); oxs.append(o.get_frame_grad(0)[0])
    if A:
        print(f'  [{label}] action grad: vs f32 %.1e vs f64 %.1e floor %.1e' % (relerr(ga, ogs[0]), relerr(ga, ogs[1]), relerr(ogs[0], ogs[1])))
        if verbose: print('    cuda', ga.ravel()); print('    o32 ', ogs[0].ravel()); print('    o64 ', ogs[1].ravel())
    d = np.abs(a - oxs[1]).max(1); d32 = np.abs(oxs[0] - oxs[1]).max(1)
    sc = np.abs(oxs[1]).max()
    print(f'  [{label}] x.grad[0]: vs f32 %.1e vs f64 %.1e floor %.1e; particles with err > 1e-4*max: cuda %d, o32 %d of %d' %
          (relerr(a, oxs[0]), relerr(a, oxs[1]), relerr(oxs[0], oxs[1]), (d > 1e-4 * sc).sum(), (d32 > 1e-4 * sc).sum(), n))
    w = np.argsort(-d)[:3] if verbose else []
    for p in w:
        print('    worst particle', p, 'x0', x0[p], 'cuda', a[p], 'o32', oxs[0][p], 'o64', oxs[1][p])


def safe(*a, **k):
    try:
        run(*a, **k)
    except Exception as e:
        import traceback; traceback.print_exc()
        print('  FAILED', k.get('label'), repr(e))


if __name__ == '__main__':
    which = sys.argv[1:] or ['rope', 'torus']
    if 'rope3' in which:
        safe('Rope-v1', label='rope full run A', verbose=False)
        safe('Rope-v1', label='rope full run B', verbose=False)
        safe('Rope-v1', label='rope full, sort off', sort=False, verbose=False)
    if 'rope2' in which:
        kw = dict(H=1, verbose=False)
        safe('Rope-v1', label='H=1 baseline', **kw)
        safe('Rope-v1', label='H=1 sort off', sort=False, **kw)
        safe('Rope-v1', label='H=1 zero actions', act_scale=0.0, **kw)
        safe('Rope-v1', label='H=1 actions x0.1', act_scale=0.07, **kw)
        safe('Rope-v1', label='H=1 spheres lifted out of contact', lift=[(0, 0.3), (1, 0.3)], **kw)
        safe('Rope-v1', label='H=1 sphere 1 only', keep=[1], **kw)
        safe('Rope-v1', label='H=1 sphere 1 + cylinder', keep=[1, 2], **kw)
        safe('Rope-v1', label='H=1 S=5', S_=5, **kw)
        safe('Rope-v1', label='H=1 S=10', S_=10, **kw)
        safe('Rope-v1', label='H=1 gravity-free-ish: ground friction 0 ', gf=0.0, **kw)
        os.environ['DSK_FORCE_BIG'] = '1'; os.environ['DSK_FLAT_GRID'] = '1'
        safe('Rope-v1', label='H=1 batched kernel family', **kw)
    if 'rope' in which:
        safe('Rope-v1', label='rope full')
        safe('Rope-v1', label='rope full, full tape', slots=3)
        safe('Rope-v1', keep=[2], label='rope cylinder only')
        safe('Rope-v1', keep=[0, 1], label='rope spheres only')
        safe('Rope-v1', keep=[], label='rope no tools')
        safe('Rope-v1', label='rope full, yield 200', ys=200.)
        safe('Rope-v1', label='rope full, ground friction 1.5', gf=1.5)
        safe('Rope-v1', label='rope full H=1', H=1)
    if 'torus' in which:
        safe('Torus-v1', label='torus')
        safe('Torus-v1', keep=[], label='torus no tools')
        safe('Torus-v1', label='torus, ground friction 1.5', gf=1.5)
