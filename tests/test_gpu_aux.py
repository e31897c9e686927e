"""Auxiliary entry points of the C ABI against the oracle: grid-mass observation (+grad), contact distances (+grad),
adjoint decay, checkpoint copies, per-particle material fields."""
import numpy as np
import pytest

from gpu_common import ENVS, actions_for, f32, make_pair
from helpers import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['CutRearrange-v1', 'LiftSpread-v1'])
def test_grid_mass_observation_and_grad(name):
    """MPMSimulator.compute_grid_m_kernel (+.grad), mpm_simulator.py:456-471."""
    scene, eng, o = make_pair(name, n=900, max_steps=1)
    m = eng.compute_grid_m(0)[0]
    om = o.compute_grid_m(0)
    assert relerr(m, om) < 1e-5
    assert m.sum() == pytest.approx(900 * scene.p_mass, rel=1e-4)
    g = f32(np.random.RandomState(0).normal(size=m.shape))
    eng.zero_grad()
    o.zero_grad()
    eng.compute_grid_m_grad(0, g[None])
    o.compute_grid_m_grad(0, g)
    assert relerr(eng.get_particle_grad(0)[0], o.get_frame_grad(0)[0]) < 1e-4


@pytest.mark.parametrize('name', ENVS)
def test_min_dist_and_grad(name):
    """GradModel.compute_min_dist (+.grad), function.py:79-88."""
    scene, eng, o = make_pair(name, n=700, max_steps=1)
    d = eng.compute_min_dist(0)[0]
    od = o.compute_min_dist(0)
    assert d.shape == od.shape and relerr(d, od) < 1e-5
    g = f32(np.random.RandomState(1).normal(size=d.shape))
    eng.zero_grad()
    o.zero_grad()
    eng.compute_min_dist_grad(0, g[None])
    o.compute_min_dist_grad(0, g)
    assert relerr(eng.get_particle_grad(0)[0], o.get_frame_grad(0)[0]) < 1e-4
    assert relerr(eng.get_tool_grads(0), o.get_tool_grads(0)) < 2e-3


def test_decay_copy_and_material_fields():
    name = 'CutRearrange-v1'
    scene, eng, o = make_pair(name, n=500, max_steps=2)
    n = 500
    rng = np.random.RandomState(2)
    gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)))
    gt = f32(rng.normal(size=(eng.K, 8)))
    gt[0, 7] = 0
    eng.zero_grad()
    eng.add_particle_grad(1, gx[None], gv[None])
    eng.add_tool_grad(1, gt[None])
    eng.scale_grad(1, 0.25)                                   # decay_kernel, function.py:66-77
    a = eng.get_particle_grad(1)
    assert np.allclose(a[0], 0.25 * gx) and np.allclose(a[1], 0.25 * gv)
    tg = eng.get_tool_grads(1)
    assert np.allclose(tg[:, :7], 0.25 * gt[:, :7]) and np.allclose(tg[:, 7], gt[:, 7])   # gap.grad is not decayed
    # copyframe (mpm_simulator.py:368-378)
    eng.copy_step(0, 2)
    for f0, f2 in zip(eng.get_particles(0), eng.get_particles(2)):
        assert np.array_equal(f0, f2)
    assert np.array_equal(eng.get_tool_states(0), eng.get_tool_states(2))
    # per-particle yield stress (sim.yield_stress field): softer half must differ from the oracle's uniform run ...
    ys = np.full(n, scene.yield_stress, np.float32)
    ys[: n // 2] = 20.0
    eng.set_material(0, yield_stress=ys)
    o.set_material(ys=ys)
    act = actions_for(scene, 1, scale=0.7)[0]
    eng.set_action(0, act[None])
    eng.forward_step(0)
    o.forward_step(0, act)
    x, v, F, C = eng.get_particles(1)
    ox, ov, oF, oC = o.get_frame(scene.substeps)
    assert relerr(x, ox) < 1e-5 and relerr(F, oF) < 1e-4          # ... and match the oracle given the same field


def test_reference_backend_probe_reports():
    """The run-time `import taichi` probe BASELINE.md section 3 / SURVEY.md section 8c promise: reports (and records in the
    parity log) whether the reference's own backend exists on the GPU box.  It never has; if it ever does, this line in the
    log is the cue to diff against plb/engine directly."""
    import sys
    sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
    from __graft_entry__ import probe_reference_backend
    from gpu_common import parity_log_path
    verdict = probe_reference_backend()
    print(verdict)
    with open(parity_log_path(), 'a') as f:
        f.write(__import__('json').dumps(dict(test='reference_backend_probe', verdict=verdict)) + '\n')
    assert 'taichi' in verdict
