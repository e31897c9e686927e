"""Diagnostic: tool adjoints CUDA vs oracle after one adjoint substep (GPU box)."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from gpu_common import *
from helpers import relerr, small_dough, perturbed_state, tool_start
from diffskill_b200.engine import Engine
from oracle import oracle as orc
np.set_printoptions(precision=5, linewidth=200)

def run(name, nopairs, f64=False, seed_tools=True):
    scene, cfg, x0 = small_dough(name, 1200)
    scene = copy.deepcopy(scene); scene.substeps = 1
    if nopairs: scene.pairs = []
    v0, F0, C0 = perturbed_state(x0, 1)
    x0, v0, F0, C0 = f32(x0), f32(v0), f32(F0), f32(C0)
    st0 = [f32(s) for s in tool_start(name, scene)]
    eng = Engine(scene, n_envs=1, capacity=1200, max_steps=2)
    eng.set_particles(0, 0, x0, v0, F0, C0)
    o = orc.Oracle(scene, 1200, 3, f64=f64, threads=1)
    o.set_frame(0, x0, v0, F0, C0)
    for i, s in enumerate(st0):
        eng.set_tool_state(0, 0, i, s); o.set_tool_state(0, i, s)
    acts = actions_for(scene, 1, scale=1.0 / 19)
    eng.set_action(0, acts[0][None]); o.set_action(0, acts[0], n_substeps=1)
    eng.substep(0); o.substep(0)
    print('  poses f+1 err', relerr(eng.get_tool_states(1), o.get_tool_states(1)), 'cidx', o.collision_idx(1))
    for i in range(eng.K): o.set_tool_state(1, i, eng.get_tool_state(1, 0, i))
    x, v, F, C = eng.get_particles(1); o.set_frame(1, x, v, F, C)
    rng = np.random.RandomState(7); n = 1200
    gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
    gF, gC = f32(rng.normal(size=(n, 3, 3)) * 0.1), f32(rng.normal(size=(n, 3, 3)) * 1e-3)
    gt = f32(rng.normal(size=(eng.K, 8)) * 0.1) * (1 if seed_tools else 0)
    for i, t in enumerate(scene.tools):
        if t.state_dim == 7: gt[i, 7] = 0
    eng.zero_grad(); o.zero_grad()
    eng.add_particle_grad(1, gx[None], gv[None], gF[None], gC[None]); eng.add_tool_grad(1, gt[None])
    o.add_frame_grad(1, gx, gv, gF, gC)
    for i in range(eng.K): o.add_tool_grad(1, i, gt[i])
    eng.substep_grad(0); o.substep_grad(0); o.L.orc_set_velocity_grad(o.h, 0, 1)
    for i in range(eng.K):
        e1 = eng.debug_tool_frame_grad(1, 0, i); e0 = eng.debug_tool_frame_grad(0, 0, i)
        o1 = o.get_tool_grad(1, i); o0 = o.get_tool_grad(0, i)
        print('  tool', i, scene.tools[i].shape)
        print('    f+1 cuda', e1); print('    f+1 orc ', o1)
        print('    f   cuda', e0); print('    f   orc ', o0)
        print('    vel grad orc', o.get_tool_vel_grad(0, i))
    print('  action cuda', eng.get_action_grad(0)[0]); print('  action orc ', o.get_action_grad(0))
    a = eng.get_particle_grad(0); b = o.get_frame_grad(0)
    print('  particle adj err', [('%.2e' % relerr(p, q)) for p, q in zip(a, b)])

for name in ['GatherMove-v1', 'LiftSpread-v1']:
    for nopairs in (False, True):
        print('==', name, 'nopairs' if nopairs else 'pairs')
        run(name, nopairs)
print('== GatherMove f64 oracle'); run('GatherMove-v1', False, f64=True)
