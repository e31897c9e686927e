"""Legacy PlasticineLab grid loss (plb/engine/losses/loss.py) of diffskill_b200.sim.losses.

CPU part: the target-SDF Jacobi sweep against a plain-loop restatement of loss.py:105-125 and the fixed-point early exit.
GPU part (device-agnostic: also runs on the emulated engine, DSK_LIB=emu): values and adjoints of the density / SDF / contact
terms against a float64 numpy restatement built on the oracle's grid-mass and contact-distance passes.
"""
import types

import numpy as np
import pytest
import torch

from diffskill_b200.sim.losses import INF, Loss, target_sdf_sweep


def loop_sweep(density, sdf_copy, nearest_copy, dx):
    """loss.py:105-125, literally (fp32 arithmetic through numpy scalars)."""
    n = density.shape[0]
    sdf = np.full((n, n, n), INF, np.float32)
    nearest = nearest_copy.copy()
    f = np.float32
    for i in range(n):
        for j in range(n):
            for k in range(n):
                gp = np.array([i, j, k], np.float32) * f(dx)
                if density[i, j, k] > 1e-4:
                    sdf[i, j, k] = 0.
                    nearest[i, j, k] = gp
                    continue
                for ox in range(-3, 3):
                    for oy in range(-3, 3):
                        for oz in range(-3, 3):
                            v = (i + ox, j + oy, k + oz)
                            if min(v) >= 0 and max(v) < n and abs(ox) + abs(oy) + abs(oz) != 0 and sdf_copy[v] < INF:
                                d = gp - nearest_copy[v]
                                dist = np.sqrt(f((d * d).sum(dtype=np.float32) + f(1e-8)))
                                if dist < sdf[i, j, k]:
                                    nearest[i, j, k] = nearest_copy[v]
                                    sdf[i, j, k] = dist
    return sdf, nearest


def test_target_sdf_sweep_matches_the_loop_restatement():
    n, dx = 7, 1 / 7
    density = np.zeros((n, n, n), np.float32)
    density[1, 2, 1] = density[5, 5, 4] = density[5, 4, 4] = 1.0
    sc = np.full((n, n, n), INF, np.float32)
    nc = np.zeros((n, n, n, 3), np.float32)
    for it in range(3):
        es, en = loop_sweep(density, sc, nc, dx)
        s, q = target_sdf_sweep(torch.from_numpy(density), torch.from_numpy(sc), torch.from_numpy(nc), dx)
        reached = es < INF
        assert np.array_equal(reached, s.numpy() < INF)
        assert np.allclose(s.numpy(), es, rtol=0, atol=1e-6), it
        # equidistant candidates may resolve differently by round-off; the chosen point must realise the distance
        ax = np.arange(n, dtype=np.float32) * np.float32(dx)
        pos = np.stack(np.meshgrid(ax, ax, ax, indexing='ij'), -1)
        far = reached & (density <= 1e-4)
        assert np.allclose(np.sqrt(((pos - q.numpy()) ** 2).sum(-1) + 1e-8)[far], es[far], atol=1e-6)
        sc, nc = es, en
    assert (es[density > 1e-4] == 0).all() and es.max() < INF


def test_update_target_stops_at_the_fixed_point_of_the_sweep():
    n = 8
    sim = types.SimpleNamespace(engine=None, n_grid=n, dx=1 / n, primitives=[], _frame_to_step=lambda f: f)
    L = Loss(None, sim, device='cpu')
    g = np.zeros((n, n, n), np.float32)
    g[3:5, 2:4, 4] = 0.7
    L.target_density = torch.from_numpy(g)
    L.update_target()
    assert L.sweeps < 2 * n
    s, q = target_sdf_sweep(L.target_density, L.target_sdf_copy, L.nearest_point_copy, L.dx)      # any further sweep: identity
    assert torch.equal(s, L.target_sdf) and torch.equal(q, L.nearest_point)
    # exact distances to the nearest solid node (3-cell reach per sweep covers the grid well before 2n sweeps)
    solid = np.argwhere(g > 1e-4) / n
    ax = np.arange(n) / n
    P = np.stack(np.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(-1, 3)
    exact = np.sqrt(((P[:, None] - solid[None]) ** 2).sum(-1) + 1e-8).min(1).reshape(n, n, n)
    exact[g > 1e-4] = 0
    assert np.allclose(L.target_sdf.numpy(), exact, atol=1e-6)


# ---- against the oracle's passes -------------------------------------------------------------------------------------------
def _expected(o, scene, tdens, tsdf, wd, ws, wc, soft, cols_of):
    m = o.compute_grid_m(0).astype(np.float64)
    cols = o.compute_min_dist(0).astype(np.float64)
    density, sdf = np.abs(m - tdens).sum(), (tsdf * m).sum()
    gm = wd * np.sign(m - tdens) + ws * tsdf
    gc = np.zeros_like(cols)
    contact, mins = 0., []
    for l, r in cols_of:
        if r - l == 2:
            first = cols[:, l] < cols[:, l + 1]
            s = np.where(first, cols[:, l], cols[:, l + 1])
            col = np.where(first, l, l + 1)
        else:
            s, col = cols[:, l], np.full(len(cols), l)
        d = np.maximum(s, 0.)
        if soft:
            w = 1 / (1 + 1e4 * d * d)
            dw = -2e4 * d * w * w
            N, S = w.sum(), (d * w).sum()
            M = S / N
            dM = (w + d * dw) / N - S * dw / N ** 2
        else:
            i = int(np.argmin(d))
            M = d[i]
            dM = np.zeros_like(d)
            dM[i] = 1.
        mins.append(M)
        contact += wc * M * M
        gs = 2 * wc * M * dM * (s > 0)
        gc[np.arange(len(cols)), col] += gs
    return density, sdf, contact, mins, gm, gc


@pytest.mark.gpu
@pytest.mark.parametrize('soft', [False, True])
@pytest.mark.parametrize('name', ['GatherMove-v1', 'LiftSpread-v1'])
def test_legacy_loss_terms_and_adjoints_match_the_oracle(name, soft):
    from gpu_common import f32, make_pair
    from helpers import relerr
    scene, eng, o = make_pair(name, n=700, max_steps=1)
    n = scene.n_grid
    sim = types.SimpleNamespace(engine=eng, n_grid=n, dx=scene.dx, primitives=scene.tools,
                                _frame_to_step=lambda f: f // scene.substeps)
    L = Loss(None, sim)
    wd, ws, wc = 10., 10., 1.
    L.set_weights_only(ws, wd, wc, soft, 0.)
    ax = np.arange(n) / n
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    r = np.sqrt((X - 0.45) ** 2 + (Y - 0.1) ** 2 + (Z - 0.5) ** 2)
    tdens, tsdf = f32((r < 0.08) * scene.p_mass * 6.), f32(np.maximum(r - 0.08, 0.))
    L.target_density = torch.from_numpy(tdens).to(L.device)
    L.target_sdf = torch.from_numpy(tsdf).to(L.device)
    L._target_iou = L._iou_of(L.target_density)
    assert L._target_iou == pytest.approx(1.0, rel=1e-5)          # I/(U-I) of a grid with itself, loss.py:68-70

    density, sdf, contact, mins, gm, gc = _expected(o, scene, tdens, tsdf, wd, ws, wc, soft, L._cols)
    L.reset()
    info = L.compute_loss(0)                                       # second evaluation of frame 0: loss accumulates (loss.py:210)
    total = contact + wd * density + ws * sdf
    assert L.density_loss == pytest.approx(density, rel=2e-5)
    assert L.sdf_loss == pytest.approx(sdf, rel=2e-5)
    assert L.contact_loss == pytest.approx(contact, rel=2e-4, abs=1e-9)
    assert np.allclose(L.min_dist, mins, rtol=2e-4, atol=1e-7)
    assert L._start_loss == pytest.approx(total, rel=2e-5)
    assert info['loss'] == pytest.approx(2 * total, rel=2e-5) and info['reward'] == pytest.approx(-total, rel=2e-4)
    m64 = o.compute_grid_m(0).astype(np.float64)
    I = (m64 * tdens).sum() / m64.max() / tdens.max()
    U = m64.sum() / m64.max() + tdens.sum() / tdens.max()
    assert info['iou'] == pytest.approx(I / (U - I), rel=1e-4)
    assert L.get_state() == {'_start_loss': L._start_loss, '_last_loss': L._last_loss, '_init_iou': L._init_iou}

    eng.zero_grad()
    o.zero_grad()
    L.compute_loss_kernel_grad(0)
    o.compute_grid_m_grad(0, f32(gm))
    o.compute_min_dist_grad(0, f32(gc))
    assert relerr(eng.get_particle_grad(0)[0], o.get_frame_grad(0)[0]) < 2e-4
    if np.abs(gc).max() > 0:
        assert relerr(eng.get_tool_grads(0), o.get_tool_grads(0)) < 2e-3


@pytest.mark.gpu
def test_grid_loss_on_the_env_classes():
    """Loss(cfg.ENV.loss, env.simulator) as the reference builds it (taichi_env.py:58; plb/envs/env.py in PlasticineLab):
    target SDF from a density grid, reward bookkeeping over env steps, and the loss adjoint reaching x.grad of the frame."""
    from diffskill_b200.config import load
    from diffskill_b200.envs.scenes import SCENES
    from diffskill_b200.sim import Loss as L2, TaichiEnv
    assert L2 is Loss
    cfg = load(data=SCENES['GatherMove-v1'])
    te = TaichiEnv(cfg, loss=False, max_env_steps=2)
    te.initialize()
    sim = te.simulator
    loss = Loss(cfg.ENV.loss, sim)
    loss.initialize()                                              # weights 10 / 10 / 1, hard contact, no target file
    assert loss.sdf_weight[None] == 10 and loss.contact_weight[0] == 1 and not loss.soft_contact_loss
    assert [p.action_dim for p in loss.primitives] == [7, 6] and loss._cols == [(0, 2), (2, 3)]
    n = sim.n_grid
    ax = np.arange(n) / n
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    blob = (np.sqrt((X - 0.5) ** 2 + (Y - 0.1) ** 2 + (Z - 0.5) ** 2) < 0.07).astype(np.float32) * float(sim.p_mass) * 8
    loss.load_target_density(grids=blob)
    assert 0 < loss.sweeps < 2 * n and loss._target_iou == pytest.approx(1.0, rel=1e-5)
    tsdf = loss.target_sdf.cpu().numpy()
    assert (tsdf[blob > 1e-4] == 0).all() and tsdf.max() < loss.inf
    far = tsdf[int(0.9 * n), int(0.1 * n), int(0.5 * n)]          # a node 0.4 from the blob centre
    assert far == pytest.approx(0.4 - 0.07, abs=1.5 / n)
    loss.reset()
    assert loss._start_loss > 0 and 0 <= loss._init_iou < 1
    te.step(np.zeros(13))
    sim.cur = 0                                                    # copy mode keeps the state at frame 0
    info = loss.compute_loss(0)
    assert set(info) >= {'loss', 'reward', 'incremental_iou', 'iou', 'target_iou', 'sdf_loss', 'density_loss', 'contact_loss'}
    assert np.isfinite(info['loss']) and info['reward'] == pytest.approx(loss._start_loss - info['loss'], rel=1e-6)
    sim.engine.zero_grad()
    loss.compute_loss_kernel_grad(0)
    g = sim.engine.get_particle_grad(0)[0]
    assert np.isfinite(g).all() and np.abs(g).max() > 0
    # the density + SDF pull: moving every particle along -grad lowers the loss (first order)
    x = sim.get_x(0)
    before = loss.density_loss * 10 + loss.sdf_loss * 10 + loss.contact_loss
    sim.reset(x - 2e-4 * g[:len(x), :3] / np.abs(g).max())
    loss.clear()
    after = loss.compute_loss(0)['loss']
    assert after < before


@pytest.mark.gpu
def test_grid_loss_of_one_env_of_a_batch():
    """A Loss bound to env 1 of a two-env engine sees that env's particles only, and its adjoint touches that env only."""
    from diffskill_b200.engine import Engine
    from diffskill_b200.scene import load_scene
    from diffskill_b200.shapes import make_box
    scene, _ = load_scene('CutRearrange-v1')
    n = 500
    xs = [make_box((0.45 + 0.1 * b, 0.07, 0.5), (0.1, 0.06, 0.08), n, np.random.RandomState(b)).astype(np.float32) for b in range(2)]
    both = Engine(scene, n_envs=2, capacity=n, max_steps=1)
    single = Engine(scene, n_envs=1, capacity=n, max_steps=1)
    for b in range(2):
        both.set_particles(0, b, xs[b])
    single.set_particles(0, 0, xs[1])
    ng = scene.n_grid
    mk = lambda eng, env: Loss(None, types.SimpleNamespace(engine=eng, n_grid=ng, dx=scene.dx, primitives=scene.tools,
                                                            _frame_to_step=lambda f: f // scene.substeps), env=env)
    La, Lb = mk(both, 1), mk(single, 0)
    rng = np.random.RandomState(0)
    tdens = torch.from_numpy((rng.uniform(size=(ng, ng, ng)) < 0.01).astype(np.float32) * scene.p_mass)
    tsdf = torch.from_numpy(rng.uniform(0, 0.3, size=(ng, ng, ng)).astype(np.float32))
    for L in (La, Lb):
        L.set_weights_only(10., 10., 1., True, 0.)
        L.target_density, L.target_sdf = tdens.to(L.device), tsdf.to(L.device)
        L.compute_loss_kernel(0)
    assert La.loss == pytest.approx(Lb.loss, rel=1e-6) and La.min_dist == pytest.approx(Lb.min_dist, rel=1e-5)
    both.zero_grad()
    single.zero_grad()
    La.compute_loss_kernel_grad(0)
    Lb.compute_loss_kernel_grad(0)
    assert np.allclose(both.get_particle_grad(0, 1)[0], single.get_particle_grad(0, 0)[0], rtol=1e-5, atol=1e-9)
    assert np.abs(both.get_particle_grad(0, 0)[0]).max() == 0
