"""Generates the committed golden fixtures from the oracle (run from the repo root: python tests/golden/make_golden.py).

PARITY UNPINNED: the reference (taichi==0.7.26) cannot be imported offline and ships no golden vectors for this
path, so these fixtures pin the ORACLE (fp64 twin) against regressions and give the CUDA path a fixed target; they
are not outputs of the reference itself.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import ENVS, perturbed_state, small_dough, tool_start  # noqa: E402
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, n=300, seed=0):
    scene, cfg, x0 = small_dough(name, n, seed)
    v0, F0, C0 = perturbed_state(x0, seed + 1, vel=0.05, strain=0.01)
    x0, v0, F0, C0 = [a.astype(np.float32) for a in (x0, v0, F0 / 1.0, C0 * 0.2)]
    st0 = np.stack([np.asarray(s, np.float32) for s in tool_start(name, scene)])
    S = scene.substeps
    act = np.random.RandomState(3).uniform(-0.7, 0.7, scene.action_dim).astype(np.float32)
    rng = np.random.RandomState(9)
    gx, gv = rng.normal(size=(n, 3)).astype(np.float32), (rng.normal(size=(n, 3)) * 0.01).astype(np.float32)
    out = dict(x0=x0, v0=v0, F0=F0, C0=C0, tools0=st0, action=act, gx=gx, gv=gv)
    o = orc.Oracle(scene, n, S + 1, f64=True, threads=1)
    o.set_frame(0, x0, v0, F0, C0)
    for i, s in enumerate(st0):
        o.set_tool_state(0, i, s)
    base, key = o.cell_index(0)
    o.forward_step(0, act)
    x, v, F, C = o.get_frame(S)
    o.zero_grad()
    o.add_frame_grad(S, gx, gv)
    ga = o.backward_step(0)
    g0 = o.get_frame_grad(0)
    out.update(base0=base, occupancy0=np.packbits(o.occupancy(0)), x1=x, v1=v, F1=F, C1=C, tools1=o.get_tool_states(S),
               action_grad=ga, gx0=g0[0], gv0=g0[1], gF0=g0[2], gC0=g0[3], tool_grad0=o.get_tool_grads(0))
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(name, 'saved', {k: getattr(v, 'shape', None) for k, v in out.items()})


if __name__ == '__main__':
    for nm in (sys.argv[1:] or ENVS):
        make(nm)
