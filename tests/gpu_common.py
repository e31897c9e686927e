"""Helpers shared by the -m gpu parity tests: build identical engine / oracle pairs."""
import copy

import numpy as np

from helpers import ENVS, perturbed_state, relerr, small_dough, tool_start
from diffskill_b200.engine import Engine
from oracle import oracle as orc



# torch device of the API-level tests: the engine's CPU emulation (DSK_LIB=emu, tests/host_check) works on host memory
import os
DEVICE = 'cpu' if os.environ.get('DSK_LIB') == 'emu' else 'cuda'


def f32(a):
    return np.asarray(a, dtype=np.float32)


def make_pair(name, n=1500, substeps=None, max_steps=4, seed=0, sort=True, step_slots=1, twin=False, grid_tape_mib=256, threads=1):
    """Engine + fp32 oracle (+ fp64 twin if twin=True) on the same fp32-representable inputs.
    substeps=1 turns every env step into one substep."""
    scene, cfg, x0 = small_dough(name, n, seed)
    scene = copy.deepcopy(scene)
    if substeps is not None:
        scene.substeps = substeps
    v0, F0, C0 = perturbed_state(x0, seed + 1)
    x0, v0, F0, C0 = f32(x0), f32(v0), f32(F0), f32(C0)
    st0 = [f32(s) for s in tool_start(name, scene)]
    eng = Engine(scene, n_envs=1, capacity=n, max_steps=max_steps, step_slots=step_slots, sort=sort,
                 grid_tape_mib=grid_tape_mib)
    eng.set_particles(0, 0, x0, v0, F0, C0)
    oracles = []
    for f64 in ([False, True] if twin else [False]):
        o = orc.Oracle(scene, n, max_steps * scene.substeps + 1, f64=f64, threads=threads)
        o.set_frame(0, x0, v0, F0, C0)
        oracles.append(o)
    for i, s in enumerate(st0):
        eng.set_tool_state(0, 0, i, s)
        for o in oracles:
            o.set_tool_state(0, i, s)
    if twin:
        return scene, eng, oracles[0], oracles[1]
    return scene, eng, oracles[0]


def within_noise_floor(err_cuda_vs_f64, err_f32_vs_f64, tol, factor=3.0):
    """The acceptance rule of the parity tests.  `tol` is the north-star tolerance.  Where the reference's own
    fp32 formulation is ill-conditioned (collider velocity = pose difference / dt, branchy contact response,
    1/clamp(sigma_j^2 - sigma_i^2) in the SVD adjoint) the fp32 oracle itself sits further than `tol` from its fp64
    twin; there the CUDA path must be no further from the fp64 twin than `factor` times the fp32 oracle is."""
    return err_cuda_vs_f64 <= max(tol, factor * err_f32_vs_f64)


def sync_oracle_to_engine(eng, o, step, f):
    """Copy the engine's checkpoint `step` into the oracle's frame f (so the next substep starts from identical bits)."""
    x, v, F, C = eng.get_particles(step)
    o.set_frame(f, x, v, F, C)
    for i in range(eng.K):
        o.set_tool_state(f, i, eng.get_tool_state(step, 0, i))


def actions_for(scene, steps, seed=3, scale=1.0):
    return f32(np.random.RandomState(seed).uniform(-1, 1, (steps, scene.action_dim)) * scale)


# ---- reviewable record of the measured parity errors --------------------------------------------------------------------------
# Every parity test appends one JSON line per (test, scene, quantity): the CUDA path's distance from the fp32 oracle and
# from its fp64 twin, the fp32 oracle's own distance from the twin, the tolerance and WHICH rule accepted the case
# ("tolerance": within the north-star tolerance of the fp32 oracle; "noise-floor": within 3x the fp32-vs-fp64 gap of the
# oracle itself; "FAIL").  On the GPU box the file lands in gpurun_out/ (the only directory that travels back); the copy
# of the round is committed as profiles/parity_r02.jsonl.
def parity_log_path():
    p = os.environ.get('DSK_PARITY_LOG')
    if p:
        return p
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    return os.path.join(out, 'parity_emu.jsonl' if os.environ.get('DSK_LIB') == 'emu' else 'parity_r02.jsonl')


def record_parity(test, scene, quantity, err_f32, err_f64=None, floor=None, tol=None, config=''):
    import json
    if tol is not None and err_f32 < tol:
        rule = 'tolerance'
    elif tol is not None and err_f64 is not None and floor is not None and within_noise_floor(err_f64, floor, tol):
        rule = 'noise-floor'
    else:
        rule = 'FAIL' if tol is not None else 'report'
    rec = dict(test=test, scene=scene, config=config, quantity=quantity, err_vs_f32_oracle=float(err_f32),
               err_vs_f64_twin=None if err_f64 is None else float(err_f64),
               f32_oracle_vs_f64_twin=None if floor is None else float(floor), tol=tol, rule=rule)
    try:
        with open(parity_log_path(), 'a') as f:
            f.write(json.dumps(rec) + '\n')
    except OSError:
        pass
    return rule
