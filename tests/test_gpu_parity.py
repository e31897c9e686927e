"""CUDA path vs CPU oracle, through the C ABI (run on the GPU box: pytest -m gpu).

Tolerances are the north-star's: integer work bit-exact; fp32 state within 1e-5 (normwise
relative) per substep; action gradients within 1e-3.
"""
import os

import numpy as np
import pytest

from gpu_common import ENVS, actions_for, f32, make_pair, record_parity, sync_oracle_to_engine, within_noise_floor
from helpers import relerr

pytestmark = pytest.mark.gpu

TOL_STATE = 1e-5
# C (and the grid velocity it is gathered from) carries the rounding noise of the order-unspecified float atomics
TOL_STATE_C = 2e-5
TOL_GRAD_SUBSTEP = 2e-4
TOL_ACTION_GRAD = 1e-3


def test_svd_matches_oracle():
    from diffskill_b200.engine import Engine
    from diffskill_b200.scene import load_scene
    from oracle import oracle as orc
    scene, _ = load_scene('CutRearrange-v1')
    eng = Engine(scene, capacity=128, max_steps=1)
    rng = np.random.RandomState(0)
    F = np.eye(3)[None] + rng.normal(size=(4096, 3, 3)) * rng.choice([1e-7, 1e-3, 0.1, 1.0], size=(4096, 1, 1))
    F[0] = np.eye(3)
    # equal column norms with off-diagonals whose squares underflow: the rotation angle is 45 degrees however tiny the
    # coupling is (round 2 regression: 1/sqrt(d^2 + 4 ga^2) = inf gave NaN velocities two substeps into a rollout)
    for i, eps in enumerate((1e-25, 1e-30, 1e-38, 3e-20), start=1):
        F[i] = np.eye(3)
        F[i][0, 1] = eps
        F[i][2, 1] = -eps
    F = f32(F)
    U, s, V = eng.debug_svd(F)
    assert np.isfinite(U).all() and np.isfinite(s).all() and np.isfinite(V).all()
    rec = np.einsum('nij,nj,nkj->nik', U, s, V)
    assert np.abs(rec - F).max() < 5e-6 * max(1.0, np.abs(F).max())
    assert np.abs(np.einsum('nji,njk->nik', U, U) - np.eye(3)).max() < 5e-6
    assert np.abs(np.einsum('nji,njk->nik', V, V) - np.eye(3)).max() < 5e-6
    assert (np.linalg.det(U.astype(np.float64)) > 0.99).all() and (np.linalg.det(V.astype(np.float64)) > 0.99).all()
    assert (s[:, 0] >= s[:, 1]).all() and (s[:, 1] >= np.abs(s[:, 2]) - 1e-6).all()
    for i in range(0, 64):
        Uo, so, Vo = orc.svd3(F[i])
        assert np.abs(so - s[i]).max() < 5e-6 * max(1.0, np.abs(so).max())


@pytest.mark.parametrize('name', ENVS)
def test_cell_index_and_sort_bit_exact(name):
    scene, eng, o = make_pair(name, n=3000, substeps=1)
    base, key = eng.debug_cell_index(0)
    ob, _ = o.cell_index(0)
    assert np.array_equal(base, ob)                      # bit-exact cell indices
    n, nt = scene.n_grid, scene.n_grid // 4
    X, Y, Z = ob[:, 0], ob[:, 1], ob[:, 2]
    okey = ((((X >> 2) * nt + (Y >> 2)) * nt + (Z >> 2)) << 6) | ((X & 3) << 4) | ((Y & 3) << 2) | (Z & 3)
    assert np.array_equal(key, okey.astype(np.int32))    # bit-exact sort keys
    eng.set_action(0, np.zeros((1, scene.action_dim), np.float32))
    eng.substep(0)
    perm = eng.debug_sort_order()
    assert np.array_equal(np.sort(perm), np.arange(len(perm)))   # a permutation
    assert (np.diff(key[perm]) >= 0).all()                       # sorted by key
    occ = eng.debug_grid(v_out=False, m=False, occupied=True)[3]
    assert np.array_equal(occ, o.occupancy(0))           # bit-exact grid occupancy


@pytest.mark.parametrize('name', ['GatherMove-v1', 'CutRearrange-v1'])
def test_sort_and_frame_permutation_batched_layout(name, monkeypatch):
    """The sort path of batched engines -- histogram scanned over flagged chunks only, permutation from k_sort_perm, frames
    moved through shared memory by k_permute_rows (gather into sorted order, scatter back to the caller's order) -- is a pure
    permutation: sorted by the same bit-exact keys, and one whole env step reproduces the single-scene path's particles in the
    CALLER's order to rounding (the two paths use different scatter variants, so not bit for bit)."""
    scene, small, o = make_pair(name, n=3000)
    monkeypatch.setenv('DSK_FORCE_BIG', '1')
    monkeypatch.setenv('DSK_FLAT_GRID', '1')
    _, big, _ = make_pair(name, n=3000)
    base, key = big.debug_cell_index(0)
    assert np.array_equal(base, o.cell_index(0)[0])
    act = actions_for(scene, 1, scale=0.7)[0][None]
    for e in (small, big):
        e.set_action(0, act)
        e.forward_step(0)
    perm = big.debug_sort_order()
    assert np.array_equal(np.sort(perm), np.arange(len(perm)))   # a permutation
    assert (np.diff(key[perm]) >= 0).all()                       # sorted by key
    for a, b, q in zip(small.get_particles(1), big.get_particles(1), 'xvFC'):
        err = relerr(b, a)
        record_parity('sort_and_frame_permutation_batched_layout', name, q, err, tol=1e-4,
                      config='n=3000, one env step, vs the single-scene path')
        assert err < 1e-4, (q, err)
    # a second step sorts again from the un-permuted checkpoint: the histogram and its chunk flags must have been cleaned up
    for e in (small, big):
        e.set_action(1, act)
        e.forward_step(1)
    assert relerr(big.get_particles(2)[0], small.get_particles(2)[0]) < 1e-4


def _fwd_errs(scene, eng, o, s):
    x, v, F, C = eng.get_particles(s + 1)
    ox, ov, oF, oC = o.get_frame(s + 1)
    _, vout, m, _ = eng.debug_grid()
    _, ovout, om = o.get_grid()
    # grid velocity of nodes that carry almost no mass is v_in/m of two tiny, cancellation-prone sums whose
    # order the reference leaves unspecified (float atomics): compare momentum everywhere and velocity on
    # nodes holding at least 1% of one particle's mass.
    heavy = om > 1e-2 * scene.p_mass
    return dict(x=relerr(x, ox), v=relerr(v, ov), F=relerr(F, oF), C=relerr(C, oC),
                grid_mom=relerr(vout * m[..., None], ovout * om[..., None]),
                grid_v=relerr(vout[heavy], ovout[heavy]), grid_m=relerr(m, om),
                tools=relerr(eng.get_tool_states(s + 1), o.get_tool_states(s + 1))), (m, om)


def _oracle_errs_fwd(scene, a, b, s):
    ax, av, aF, aC = a.get_frame(s + 1)
    bx, bv, bF, bC = b.get_frame(s + 1)
    _, avout, am = a.get_grid()
    _, bvout, bm = b.get_grid()
    heavy = bm > 1e-2 * scene.p_mass
    return dict(x=relerr(ax, bx), v=relerr(av, bv), F=relerr(aF, bF), C=relerr(aC, bC),
                grid_mom=relerr(avout * am[..., None], bvout * bm[..., None]),
                grid_v=relerr(avout[heavy], bvout[heavy]), grid_m=relerr(am, bm),
                tools=relerr(a.get_tool_states(s + 1), b.get_tool_states(s + 1)))


@pytest.mark.parametrize('name', ENVS)
@pytest.mark.parametrize('sort', [True, False])
def test_substep_forward_parity(name, sort):
    steps = 6
    scene, eng, o, o64 = make_pair(name, n=1500, substeps=1, max_steps=steps, sort=sort, twin=True)
    acts = actions_for(scene, steps, scale=1.0 / 19)
    worst, worst64, floor = {}, {}, {}
    for s in range(steps):
        eng.set_action(s, acts[s][None])
        eng.substep(s)
        for oo in (o, o64):
            oo.set_action(s, acts[s], n_substeps=1)
            oo.substep(s)
        e32, (m, om) = _fwd_errs(scene, eng, o, s)
        e64, _ = _fwd_errs(scene, eng, o64, s)
        fl = _oracle_errs_fwd(scene, o, o64, s)
        for k_ in e32:
            worst[k_] = max(worst.get(k_, 0), e32[k_])
            worst64[k_] = max(worst64.get(k_, 0), e64[k_])
            floor[k_] = max(floor.get(k_, 0), fl[k_])
        assert np.array_equal(m > 1e-12, om > 1e-12), "grid occupancy predicate differs"
        for oo in (o, o64):
            sync_oracle_to_engine(eng, oo, s + 1, s + 1)
    fmt = lambda d: {k_: '%.1e' % e_ for k_, e_ in d.items()}
    print(name, 'sort' if sort else 'nosort', 'vs_f32', fmt(worst), 'vs_f64', fmt(worst64), 'f32_vs_f64', fmt(floor))
    for k_ in worst:
        tol = TOL_STATE_C if k_ in ('C', 'grid_v') else TOL_STATE
        record_parity('substep_forward', name, k_, worst[k_], worst64[k_], floor[k_], tol, config='sort' if sort else 'nosort')
        assert worst[k_] < tol or within_noise_floor(worst64[k_], floor[k_], tol), (k_, worst[k_], worst64[k_], floor[k_])


@pytest.mark.parametrize('name', ENVS)
def test_substep_backward_parity(name):
    steps = 4
    scene, eng, o, o64 = make_pair(name, n=1200, substeps=1, max_steps=steps, twin=True)
    acts = actions_for(scene, steps, scale=1.0 / 19)
    n = eng.n_particles()
    rng = np.random.RandomState(7)
    for s in range(steps):
        eng.set_action(s, acts[s][None])
        eng.substep(s)
        for oo in (o, o64):
            oo.set_action(s, acts[s], n_substeps=1)
            oo.substep(s)
            sync_oracle_to_engine(eng, oo, s + 1, s + 1)
    worst, worst64, floor = {}, {}, {}

    def collect(oo, s):
        b = oo.get_frame_grad(s)
        oga_in, _, ogm = oo.get_grid_grad()
        return dict(gx=b[0], gv=b[1], gF=b[2], gC=b[3], grid_gv=oga_in, grid_gm=ogm, tool=oo.get_tool_grads(s),
                    action=oo.get_action_grad(s))

    for s in range(steps - 1, -1, -1):
        # fresh random incoming adjoints at frame s+1 on all sides
        eng.zero_grad()
        gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
        gF, gC = f32(rng.normal(size=(n, 3, 3)) * 0.1), f32(rng.normal(size=(n, 3, 3)) * 1e-3)
        gt = f32(rng.normal(size=(eng.K, 8)) * 0.1)
        for i, t in enumerate(scene.tools):
            if t.state_dim == 7:
                gt[i, 7] = 0
        eng.add_particle_grad(s + 1, gx[None], gv[None], gF[None], gC[None])
        eng.add_tool_grad(s + 1, gt[None])
        eng.substep_grad(s)
        a = eng.get_particle_grad(s)
        ga_in, gm = eng.debug_grid_grad()
        mine = dict(gx=a[0], gv=a[1], gF=a[2], gC=a[3], grid_gv=ga_in, grid_gm=gm, tool=eng.get_tool_grads(s),
                    action=eng.get_action_grad(s)[0])
        res = []
        for oo in (o, o64):
            oo.zero_grad()
            oo.add_frame_grad(s + 1, gx, gv, gF, gC)
            for i in range(eng.K):
                oo.add_tool_grad(s + 1, i, gt[i])
            oo.substep_grad(s)
            oo.L.orc_set_velocity_grad(oo.h, s, 1)
            res.append(collect(oo, s))
        # A particle sitting exactly on the yield surface (|delta_gamma| ~ 1e-8, mpm_simulator.py:176-178) takes the
        # other branch under a 1-ulp state difference and has a different (equally valid) gradient: exclude them.
        sig = o64.get_svd()[2]
        eps_ = np.log(np.maximum(sig, 0.05))
        eh = eps_ - eps_.mean(1, keepdims=True)
        dgamma = np.sqrt((eh ** 2).sum(1) + 1e-8) - scene.yield_stress / (2 * scene.mu)
        keep = np.abs(dgamma) > 1e-5
        # (soft doughs -- yield_stress 50 in the legacy PlasticineLab scenes -- keep more particles on the surface)
        assert keep.mean() > 0.97
        for k_ in mine:
            sel = keep if k_ in ('gx', 'gv', 'gF', 'gC') else slice(None)
            worst[k_] = max(worst.get(k_, 0), relerr(mine[k_][sel], res[0][k_][sel]))
            worst64[k_] = max(worst64.get(k_, 0), relerr(mine[k_][sel], res[1][k_][sel]))
            floor[k_] = max(floor.get(k_, 0), relerr(res[0][k_][sel], res[1][k_][sel]))
    fmt = lambda d: {k_: '%.1e' % e_ for k_, e_ in d.items()}
    print(name, 'vs_f32', fmt(worst), 'vs_f64', fmt(worst64), 'f32_vs_f64', fmt(floor))
    for k_ in worst:
        record_parity('substep_backward', name, k_, worst[k_], worst64[k_], floor[k_], TOL_GRAD_SUBSTEP)
        assert worst[k_] < TOL_GRAD_SUBSTEP or within_noise_floor(worst64[k_], floor[k_], TOL_GRAD_SUBSTEP), \
            (k_, worst[k_], worst64[k_], floor[k_])


# Rope-v1 (two Spheres + a static Cylinder on a sliding ground), DESIGN.md section 10: with the first builds of round 1 the
# 3-step gradient was 3.5e-3 off on the B200 -- a ~1e-8 bias in the cosine of the Jacobi rotations (MUFU rsqrt) put the GPU
# into the scene's other gradient "basin".  The cosine has been unbiased since (svd3.cuh: rsqrt_unbiased); these cases
# passed on the B200 in the round-1 driver run and in every round-2 run (profiles/parity_r02.jsonl), so they gate again.
MULTI_STEP_ENVS = list(ENVS)


@pytest.mark.parametrize('name', MULTI_STEP_ENVS)
@pytest.mark.parametrize('slots,tape_mib', [(1, 256), (3, 256), (3, 1), (3, 0)])
def test_multi_step_action_gradient(name, slots, tape_mib):
    _multi_step_action_gradient(name, slots, tape_mib)


@pytest.mark.parametrize('name', MULTI_STEP_ENVS)
@pytest.mark.parametrize('slots,tape_mib', [(3, 256), (1, 0)])
def test_multi_step_action_gradient_batched_layout(name, slots, tape_mib, monkeypatch):
    """Same property with the kernel variants batched engines select (>= 24576 particle slots): throughput layout
    of the grid kernels and the high-occupancy particle kernels."""
    monkeypatch.setenv('DSK_FLAT_GRID', '1')
    monkeypatch.setenv('DSK_FORCE_BIG', '1')
    _multi_step_action_gradient(name, slots, tape_mib)


def _multi_step_action_gradient(name, slots, tape_mib):
    """3 env steps of the real substep count: checkpoint + recompute (slots=1) and full tape (slots=3)
    must both match the oracle's taped gradient (the reference's own property test, long_term_gradient.ipynb).
    tape_mib: grid tape on (256), overflowing -> device-side fallback to recompute (1), off (0)."""
    H = 3
    scene, eng, o, o64 = make_pair(name, n=800, max_steps=H, step_slots=slots, grid_tape_mib=tape_mib, twin=True)
    acts = actions_for(scene, H, scale=0.7)
    n = eng.n_particles()
    S = scene.substeps
    for s in range(H):
        eng.set_action(s, acts[s][None])
        eng.forward_step(s)
        o.forward_step(s, acts[s])
    x, v, F, C = eng.get_particles(H)
    ox, ov, oF, oC = o.get_frame(H * S)
    print(name, 'state after %d substeps: x %.2e v %.2e F %.2e C %.2e' % (H * S, relerr(x, ox), relerr(v, ov), relerr(F, oF), relerr(C, oC)))
    assert relerr(x, ox) < 1e-4 and relerr(v, ov) < 2e-3
    rng = np.random.RandomState(11)
    gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
    eng.zero_grad()
    o.zero_grad()
    eng.add_particle_grad(H, gx[None], gv[None])
    o.add_frame_grad(H * S, gx, gv)
    ga, oga = np.zeros((H, scene.action_dim)), np.zeros((H, scene.action_dim))
    for s in range(H - 1, -1, -1):
        eng.backward_step(s)
        ga[s] = eng.get_action_grad(s)[0]
        oga[s] = o.backward_step(s)
    e = relerr(ga, oga)
    a = eng.get_particle_grad(0)
    b = o.get_frame_grad(0)
    ex = relerr(a[0], b[0])
    print(name, 'slots', slots, 'action grad err %.2e' % e, 'x.grad[0] err %.2e' % ex)
    cfg_ = f'H=3 slots={slots} tape_mib={tape_mib}' + (' batched-layout' if os.environ.get('DSK_FORCE_BIG') == '1' else '')
    if e < TOL_ACTION_GRAD and ex < 5 * TOL_ACTION_GRAD:
        record_parity('multi_step_gradient', name, 'action_grad', e, tol=TOL_ACTION_GRAD, config=cfg_)
        record_parity('multi_step_gradient', name, 'x_grad0', ex, tol=5 * TOL_ACTION_GRAD, config=cfg_)
        return
    # Out of the north-star tolerance against the fp32 oracle: acceptable only where the reference's own fp32 formulation
    # is that far from its fp64 twin on this scene (Torus-v1: fp32 oracle 1.5e-3 from fp64, CUDA 7e-5 from fp64).
    for s in range(H):
        o64.forward_step(s, acts[s])
    o64.zero_grad()
    o64.add_frame_grad(H * S, gx, gv)
    oga64 = np.zeros((H, scene.action_dim))
    for s in range(H - 1, -1, -1):
        oga64[s] = o64.backward_step(s)
    b64 = o64.get_frame_grad(0)
    e64, f64_ = relerr(ga, oga64), relerr(oga, oga64)
    ex64, fx64 = relerr(a[0], b64[0]), relerr(b[0], b64[0])
    print(name, 'vs fp64 twin: action grad %.2e (fp32 oracle %.2e), x.grad[0] %.2e (fp32 oracle %.2e)' % (e64, f64_, ex64, fx64))
    record_parity('multi_step_gradient', name, 'action_grad', e, e64, f64_, TOL_ACTION_GRAD, config=cfg_)
    record_parity('multi_step_gradient', name, 'x_grad0', ex, ex64, fx64, 5 * TOL_ACTION_GRAD, config=cfg_)
    assert within_noise_floor(e64, f64_, TOL_ACTION_GRAD), (e, e64, f64_)
    assert within_noise_floor(ex64, fx64, 5 * TOL_ACTION_GRAD), (ex, ex64, fx64)


@pytest.mark.parametrize('name', ['LiftSpread-v1', 'GatherMove-v1', 'CutRearrange-v1'])
def test_fine_grained_substeps_equal_whole_step(name):
    """MPMSimulator.substep / substep_grad called substep by substep (mpm_simulator.py:307-345, the reference's own entry
    points) must reproduce forward_step / backward_step with substeps > 1: states, particle adjoints, action gradients.
    (Round 1 ran every fine-grained substep of a step under one tile epoch: grid_op skipped most tiles.)"""
    S = 4
    H = 2
    res = []
    for fine in (False, True):
        scene, eng, o = make_pair(name, n=900, substeps=S, max_steps=H)
        acts = actions_for(scene, H, scale=0.7)
        n = eng.n_particles()
        for s in range(H):
            eng.set_action(s, acts[s][None])
            if fine:
                for j in range(S):
                    eng.substep(s * S + j)
            else:
                eng.forward_step(s)
        state = eng.get_particles(H)
        rng = np.random.RandomState(5)
        gx, gv = f32(rng.normal(size=(n, 3))), f32(rng.normal(size=(n, 3)) * 0.01)
        eng.zero_grad()
        eng.add_particle_grad(H, gx[None], gv[None])
        for s in range(H - 1, -1, -1):
            if fine:
                for j in range(S - 1, -1, -1):
                    eng.substep_grad(s * S + j)
            else:
                eng.backward_step(s)
        res.append((state, eng.get_particle_grad(0), eng.get_action_grads(0, H)[:, 0]))
        if not fine:   # and the whole-step path against the oracle, so "equal" is not "equally wrong"
            for s in range(H):
                o.forward_step(s, acts[s])
            ox, ov, _, _ = o.get_frame(H * S)
            assert relerr(state[0], ox) < 1e-5 and relerr(state[1], ov) < 2e-3
    (sa, ga, aa), (sb, gb, ab) = res
    errs = dict(x=relerr(sb[0], sa[0]), v=relerr(sb[1], sa[1]), F=relerr(sb[2], sa[2]), C=relerr(sb[3], sa[3]),
                gx=relerr(gb[0], ga[0]), gv=relerr(gb[1], ga[1]), gF=relerr(gb[2], ga[2]), action=relerr(ab, aa))
    print(name, 'fine-grained vs whole-step', {k_: '%.1e' % e_ for k_, e_ in errs.items()})
    # different scatter orders (float atomics) and recompute vs tape only
    for k_, e_ in errs.items():
        assert e_ < (1e-3 if k_ in ('gx', 'gv', 'gF', 'action') else 2e-5), (k_, e_)


@pytest.mark.parametrize('name', ENVS)
def test_against_committed_golden_fixture(name):
    """CUDA path vs tests/golden/*.npz (oracle fp64, generated by tests/golden/make_golden.py): one env step of
    the real substep count forward + backward."""
    import os
    from helpers import small_dough
    from diffskill_b200.engine import Engine
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', f'{name}.npz'))
    n = len(g['x0'])
    scene, _, _ = small_dough(name, n)
    eng = Engine(scene, capacity=n, max_steps=1)
    eng.set_particles(0, 0, g['x0'], g['v0'], g['F0'], g['C0'])
    for i, s in enumerate(g['tools0']):
        eng.set_tool_state(0, 0, i, s)
    base, _ = eng.debug_cell_index(0)
    assert np.array_equal(base, g['base0'])
    eng.set_action(0, g['action'][None])
    eng.forward_step(0)
    x, v, F, C = eng.get_particles(1)
    assert relerr(x, g['x1']) < 1e-5 and relerr(F, g['F1']) < 1e-4 and relerr(v, g['v1']) < 2e-3
    assert relerr(eng.get_tool_states(1), g['tools1']) < 1e-6
    eng.zero_grad()
    eng.add_particle_grad(1, f32(g['gx'])[None], f32(g['gv'])[None])
    eng.backward_step(0)
    ga = eng.get_action_grad(0)[0]
    e = relerr(ga, g['action_grad'])
    print(name, 'golden: x %.1e v %.1e action grad %.1e' % (relerr(x, g['x1']), relerr(v, g['v1']), e))
    assert e < 2e-2          # same bound the fp32 oracle is held to against its fp64 twin (tests/test_oracle.py)


def test_batched_envs_ragged_and_empty():
    """Three envs in one engine with different particle counts (one of them EMPTY) must each match a single-env
    oracle run: forward step, adjoint step, per-env action gradients."""
    import copy
    from helpers import perturbed_state, small_dough, tool_start
    from diffskill_b200.engine import Engine
    from oracle import oracle as orc
    name = 'GatherMove-v1'
    counts = [700, 0, 333]
    scene, _, _ = small_dough(name, 10)
    scene = copy.deepcopy(scene)
    cap = max(counts)
    eng = Engine(scene, n_envs=3, capacity=cap, max_steps=1)
    S = scene.substeps
    acts = f32(np.random.RandomState(5).uniform(-0.7, 0.7, (3, scene.action_dim)))
    oracles, seeds = [], []
    gx = np.zeros((3, cap, 3), np.float32)
    for b, n in enumerate(counts):
        st0 = [f32(s) for s in tool_start(name, scene)]
        st0[1][0] += 0.01 * b
        for i, s in enumerate(st0):
            eng.set_tool_state(0, b, i, s)
        if n == 0:
            eng.set_particles(0, b, np.zeros((0, 3), np.float32))
            oracles.append(None)
            continue
        _, _, x0 = small_dough(name, n, seed=b)
        v0, F0, C0 = perturbed_state(x0, b + 1, vel=0.05, strain=0.01)
        x0, v0, F0, C0 = f32(x0), f32(v0), f32(F0), f32(C0 * 0.2)
        eng.set_particles(0, b, x0, v0, F0, C0)
        o = orc.Oracle(scene, n, S + 1, f64=False, threads=1)
        o.set_frame(0, x0, v0, F0, C0)
        for i, s in enumerate(st0):
            o.set_tool_state(0, i, s)
        oracles.append(o)
        gx[b, :n] = np.random.RandomState(20 + b).normal(size=(n, 3))
    assert [eng.n_particles(b) for b in range(3)] == counts
    eng.set_action(0, acts)
    eng.forward_step(0)
    eng.zero_grad()
    eng.add_particle_grad(1, gx)
    eng.backward_step(0)
    ga = eng.get_action_grad(0)
    for b, n in enumerate(counts):
        o = oracles[b]
        if o is None:
            assert np.all(ga[b] == 0) or np.isfinite(ga[b]).all()
            ts = eng.get_tool_states(1, b)
            assert np.isfinite(ts).all()
            continue
        o.forward_step(0, acts[b])
        x, v = eng.get_particles(1, b, 'xv')
        ox, ov, _, _ = o.get_frame(S)
        assert relerr(x, ox) < 1e-5 and relerr(v, ov) < 2e-3, (b, relerr(x, ox), relerr(v, ov))
        o.zero_grad()
        o.add_frame_grad(S, gx[b, :n])
        og = o.backward_step(0)
        assert relerr(ga[b], og) < 1e-3, (b, relerr(ga[b], og))


@pytest.mark.parametrize('name,tool,start', [('LiftSpread-v1', 1, (0.58, 0.04, 0.5)), ('GatherMove-v1', 1, (0.585, 0.03, 0.5))])
def test_tool_tool_collision_projection(name, tool, start):
    """The lifter starts 2 cm inside the obstacle box: set_surface_points / set_collision_idx /
    apply_collision_projection (mpm_simulator.py:286-305) must fire and match the oracle -- collision indices
    bit-exact (deterministic first minimum), poses, and the adjoint through the projection."""
    from helpers import small_dough
    scene, eng, o = make_pair(name, n=300, max_steps=2)
    S = scene.substeps
    st = np.array(scene.tools[tool].init_state, np.float32)
    st[:3] = start
    eng.set_tool_state(0, 0, tool, st)
    o.set_tool_state(0, tool, st)
    acts = actions_for(scene, 2, scale=0.5)
    hits = 0
    for s in range(2):
        eng.set_action(s, acts[s][None])
        eng.forward_step(s)
        o.forward_step(s, acts[s])
        for f in range(s * S + 1, (s + 1) * S + 1):
            ci = o.collision_idx(f)
            hits += int((ci >= 0).sum())
    assert hits > 0, "scenario must trigger the projection"
    assert relerr(eng.get_tool_states(2), o.get_tool_states(2 * S)) < 1e-5
    # collision indices of the last simulated step are still in the slot ring
    for f in range(S + 1, 2 * S + 1):
        _, ci = eng.debug_tool_frame(f, 0, tool)
        assert np.array_equal(ci, o.collision_idx(f)), f
    rng = np.random.RandomState(3)
    gt = f32(rng.normal(size=(eng.K, 8)))
    for i, t in enumerate(scene.tools):
        if t.state_dim == 7:
            gt[i, 7] = 0
    eng.zero_grad()
    o.zero_grad()
    eng.add_tool_grad(2, gt[None])
    for i in range(eng.K):
        o.add_tool_grad(2 * S, i, gt[i])
    ga, oga = np.zeros((2, scene.action_dim)), np.zeros((2, scene.action_dim))
    for s in (1, 0):
        eng.backward_step(s)
        ga[s] = eng.get_action_grad(s)[0]
        oga[s] = o.backward_step(s)
    print(name, 'projection hits', hits, 'action-grad err %.2e' % relerr(ga, oga), 'tool-grad err %.2e' % relerr(eng.get_tool_grads(0), o.get_tool_grads(0)))
    assert relerr(ga, oga) < 1e-3
    assert relerr(eng.get_tool_grads(0), o.get_tool_grads(0)) < 1e-3
