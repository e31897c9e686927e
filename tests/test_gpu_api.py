"""The reference-shaped Python API (MPMSimulator / Primitives / TaichiEnv / GradModel) over the C ABI, on the GPU,
checked against the oracle driven through the reference's own call sequences."""
import numpy as np
import pytest

from gpu_common import DEVICE
from helpers import relerr

pytestmark = pytest.mark.gpu


def _oracle_for(env, frames):
    from oracle import oracle as orc
    sim = env.simulator
    st = env.get_state()['state']
    o = orc.Oracle(sim.scene, len(st[0]), frames, f64=False, threads=8)
    o.set_frame(0, *[np.asarray(a, np.float32) for a in st[:4]])
    for i, s in enumerate(st[4:]):
        o.set_tool_state(0, i, np.asarray(s, np.float32))
    return o


@pytest.mark.parametrize('name', ['LiftSpread-v1', 'CutRearrange-v1'])
def test_random_rollout_copy_mode(name):
    """configs[0]: scripts/random_env.py -- env.step(action_space.sample()) in copy mode."""
    from diffskill_b200.envs import make
    env = make(name, seed=100)
    te = env.taichi_env
    env.reset()
    o = _oracle_for(te, te.simulator.substeps + 1)
    for t in range(3):
        a = env.sample_action()
        env.step(a)
        o.step_copy(a)
        x, v = te.simulator.get_x(0), te.simulator.get_v(0)
        ox, ov, _, _ = o.get_frame(0)
        assert relerr(x, ox) < 2e-6 * (t + 1), (t, relerr(x, ox))
        assert relerr(v, ov) < 5e-4, (t, relerr(v, ov))
        ts = np.stack([p.get_state(0)[:7] for p in te.primitives])
        assert relerr(ts, o.get_tool_states(0)[:, :7]) < 1e-6
    st = te.get_state()
    assert st['state'][0].dtype == np.float64 and st['state'][2].shape[1:] == (3, 3)   # mpm_simulator.py:382-385
    assert te.simulator.cur == 0


@pytest.mark.parametrize('name,return_dist', [('GatherMove-v1', True), ('LiftSpread-v1', False)])
def test_gradmodel_autograd_matches_oracle(name, return_dist):
    """Solver.solve_one_plan pattern (plb/optimizer/solver.py:111-127): reset, H x forward.apply, torch loss,
    backward -- the action gradient must match the oracle's taped gradient."""
    import torch
    from diffskill_b200.config import load
    from diffskill_b200.envs.scenes import SCENES
    from diffskill_b200.sim import GradModel, TaichiEnv
    cfg = load(data=SCENES[name])
    if name == 'LiftSpread-v1':
        cfg.SHAPES[0]['radius'] = 0.02          # 1005 particles: the oracle's tape AD stays fast
    te = TaichiEnv(cfg, loss=False, return_dist=return_dist, max_env_steps=4)
    te.initialize()
    n = te.n_particles
    func = GradModel(te, softness=666., return_dist=return_dist)
    H, S, A = 3, te.simulator.substeps, te.primitives.action_dim
    o = _oracle_for(te, H * S + 1)
    rng = np.random.RandomState(0)
    acts = torch.tensor(rng.uniform(-1, 1, (H, A)), device=DEVICE, dtype=torch.float32, requires_grad=True)
    ncol = 6 + (func.eng.ncols if return_dist else 0)
    W = torch.tensor(rng.normal(size=(H, n, ncol)), device=DEVICE, dtype=torch.float32)
    Wc = torch.tensor(rng.normal(size=(H, len(te.primitives), 8)) * 0.1, device=DEVICE, dtype=torch.float32)
    obs = func.reset(device=DEVICE)
    assert obs[0].shape == (n, ncol) and obs[1].shape == (len(te.primitives), 8)
    loss = 0
    for s in range(H):
        obs = func.forward(s, acts[s], *obs)
        loss = loss + (obs[0] * W[s]).sum() + (obs[1] * Wc[s]).sum()
    loss.backward()
    g = acts.grad.cpu().numpy()
    # oracle: same call sequence on the tape
    o.zero_grad()
    an = acts.detach().cpu().numpy().astype(np.float64)
    for s in range(H):
        o.forward_step(s, an[s])
    Wn, Wcn = W.cpu().numpy().astype(np.float64), Wc.cpu().numpy().astype(np.float64)
    og = np.zeros((H, A))
    for s in range(H - 1, -1, -1):
        f = (s + 1) * S
        if return_dist:
            o.compute_min_dist_grad(f, Wn[s][:, 6:])
        o.add_frame_grad(f, gx=Wn[s][:, :3], gv=Wn[s][:, 3:6])
        for i in range(len(te.primitives)):
            o.add_tool_grad(f, i, Wcn[s][i])
        og[s] = o.backward_step(s)
    e = relerr(g, og)
    print(name, 'GradModel action-grad err %.2e' % e, 'loss', float(loss))
    assert e < 1e-3


def test_field_proxies_and_errors():
    from diffskill_b200.config import load
    from diffskill_b200.envs.scenes import SCENES
    from diffskill_b200.engine import EngineError
    from diffskill_b200.sim import TaichiEnv
    cfg = load(data=SCENES['CutRearrange-v1'])
    te = TaichiEnv(cfg, loss=False, max_env_steps=2)
    te.initialize()
    sim, prims = te.simulator, te.primitives
    assert sim.n_particles[None] == 5000 and sim.substeps == 24 and sim.n_grid == 80
    prims[1].friction[None] = 3.5
    assert sim.engine.get_tool_param(1, 0) == pytest.approx(3.5)
    prims.set_softness(123.)
    assert prims.get_softness() == 123. and sim.engine.get_tool_param(0, 1) == pytest.approx(123.)
    prims[0].xyz_limit[1] = (0.9, 0.8, 0.7)
    assert sim.engine.get_tool_param(0, 6) == pytest.approx(0.8)
    sim.yield_stress.fill(77.)
    assert sim.yield_stress[0] == 77.
    assert prims[1].gap[0] == pytest.approx(0.18)
    with pytest.raises(AssertionError):
        prims[1].set_state(0, [0.5, 0.1, 0.5, 1, 0, 0, 0])          # gripper needs 8 values (primitives.py:551)
    with pytest.raises(AssertionError):
        prims.set_action(0, sim.substeps, np.zeros(3))               # wrong action length (primitives.py:865)
    prims[0].set_state(0, [0.4, 0.3, 0.5])                           # short state pads (primive_base.py:188-191)
    assert np.allclose(prims[0].get_state(0)[:3], [0.4, 0.3, 0.5]) and prims[0].get_state(0)[3] == pytest.approx(1.0)
    te.step(np.zeros(10))
    with pytest.raises(IndexError):
        sim.get_x(5)                                                   # not a step boundary and not resident
    with pytest.raises(EngineError):
        sim.engine.forward_step(7, 8, 7)                               # beyond the horizon
    sim.x.grad.fill(0)
    with pytest.raises(NotImplementedError):
        te.render()


def test_batched_solver_reduces_loss():
    """The planner loop (solver.py:97-152 pattern) over a 4-env batch: Adam on actions lowers the L2-target loss."""
    from diffskill_b200.engine import Engine
    from diffskill_b200.planner import BatchedSolver
    from diffskill_b200.scene import load_scene
    from diffskill_b200.shapes import make_box
    scene, cfg = load_scene('CutRearrange-v1')
    B, H, n = 4, 4, 600
    eng = Engine(scene, n_envs=B, capacity=n, max_steps=H, step_slots=H)
    tgt = np.zeros((B, n, 3), np.float32)
    for b in range(B):
        x = make_box((0.5, 0.06, 0.5), (0.12, 0.06, 0.06), n, np.random.RandomState(b)).astype(np.float32)
        eng.set_particles(0, b, x)
        eng.set_tool_state(0, b, 1, [0.5, 0.09, 0.5, 0.707, 0.0, 0.707, 0.0, 0.10])   # gripper around the slab
        tgt[b] = x + np.array([0.02, 0.0, 0.0], np.float32)
    solver = BatchedSolver(eng, H, lr=0.05)
    res = solver.solve(np.zeros((H, B, scene.action_dim), np.float32), solver.l2_target_loss(tgt), max_iter=8)
    h = res['history']
    print('planner loss history', ['%.3e' % v for v in h])
    assert np.isfinite(h).all() and h[-1] < h[0]
    assert res['best_action'].shape == (H, B, scene.action_dim) and np.abs(res['best_action']).max() <= 1.0


def test_multi_step_calls_equal_per_step_calls():
    """dsk_set_actions / forward_steps / loss_add_l2_steps / backward_steps issue exactly the per-step work."""
    from diffskill_b200.engine import Engine
    from diffskill_b200.scene import load_scene
    from helpers import relerr, small_dough, tool_start
    name, n, H, B = 'GatherMove-v1', 700, 3, 2
    scene, cfg, x = small_dough(name, n)
    acts = np.random.RandomState(4).uniform(-0.8, 0.8, (H, B, scene.action_dim)).astype(np.float32)
    tgt = np.broadcast_to((x + np.array([0.01, 0.0, 0.01]))[None], (B, n, 3)).astype(np.float32).copy()
    out = []
    for multi in (False, True):
        eng = Engine(scene, n_envs=B, capacity=n, max_steps=H, step_slots=H)
        for b in range(B):
            eng.set_particles(0, b, x.astype(np.float32))
            for i, st in enumerate(tool_start(name, scene)):
                eng.set_tool_state(0, b, i, np.asarray(st, np.float32))
        eng.zero_grad(); eng.loss_reset()
        if multi:
            eng.set_actions(0, acts)
            eng.forward_steps(0, H)
            eng.loss_add_l2_steps(1, H, tgt, 1.0 / H)
            eng.backward_steps(H - 1, H)
        else:
            for s in range(H):
                eng.set_action(s, acts[s]); eng.forward_step(s); eng.loss_add_l2(s + 1, tgt, 1.0 / H)
            for s in range(H - 1, -1, -1):
                eng.backward_step(s)
        out.append((eng.loss_get().copy(), eng.get_action_grads(0, H).copy(), eng.get_particles(H, 1, 'x')[0]))
    (l0, g0, x0), (l1, g1, x1) = out
    print('multi-step vs per-step: loss %.2e grads %.2e x %.2e' % (relerr(l1, l0), relerr(g1, g0), relerr(x1, x0)))
    assert np.abs(g0).max() > 0
    assert relerr(l1, l0) < 1e-5 and relerr(x1, x0) < 1e-6 and relerr(g1, g0) < 1e-4


def test_reference_solver_api():
    """plb/optimizer/solver.py: Solver(args, env).solve(init_actions, env.compute_loss, action_mask, lr, max_iter) -- torch Adam
    through GradModel's autograd Function, EMD + contact + velocity loss, clamp / mask / best-so-far, state restored."""
    import types
    import torch
    from diffskill_b200.config import load
    from diffskill_b200.envs.scenes import SCENES
    from diffskill_b200.planner import FUNCS, Solver
    from diffskill_b200.sim import TaichiEnv
    cfg = load(data=SCENES['GatherMove-v1'])
    te = TaichiEnv(cfg, loss=True, return_dist=True, max_env_steps=2)
    te.initialize()
    te.device = DEVICE
    x0 = te.simulator.get_x(0)
    te.target_x = (x0 + np.array([0.03, 0.0, 0.0])).astype(np.float32)
    te.tensor_target_x = torch.as_tensor(te.target_x, device=DEVICE)
    te.set_contact_loss_mask(torch.ones(len(te.primitives), device=DEVICE))
    args = types.SimpleNamespace(adam_loss_type='emd', stop_action_n=0, vel_loss_weight=0.02, energy_weight=0.,
                                 component_matching=False, enumerate_contact=False, debug=False)
    solver = Solver(args, te, return_dist=True, device=DEVICE)
    assert FUNCS[te] is solver.func
    H, A = 2, te.primitives.action_dim
    mask = np.ones(A, np.float32)
    mask[-1] = 0.                                   # freeze one action dimension
    np.random.seed(0)                               # compute_loss subsamples 500 particles (taichi_env.py:261)
    before = te.get_state()
    info, buffer = solver.solve(np.full((H, A), 0.2, np.float32), te.compute_loss, action_mask=mask, lr=0.05, max_iter=3,
                                verbose=False)
    assert len(buffer) == 3 and all(np.isfinite(b['loss']) for b in buffer)
    assert info['best_loss'] == min(b['loss'] for b in buffer) and info['last_loss'] == buffer[-1]['loss']
    assert info['best_action'].shape == (H, A) and np.abs(info['last_action']).max() <= 1.0
    assert (info['last_action'][:, -1] == 0).all()
    assert not np.allclose(info['last_action'][:, :-1], 0.2)            # Adam moved the free dimensions
    after = te.get_state()
    assert te.simulator.cur == 0 and np.array_equal(after['state'][0], before['state'][0])
    outs = solver.eval(info['best_action'], lambda: te.simulator.get_x(0).mean(0))
    assert len(outs) == H and np.isfinite(outs).all()


def test_cut_solve_func_api():
    """plb/cut/solve_func.solve: per-step loss with logged terms, action_dims projection, resumable optimiser state."""
    import torch
    from diffskill_b200.config import load
    from diffskill_b200.envs.scenes import SCENES
    from diffskill_b200.planner import solve
    from diffskill_b200.sim import GradModel, TaichiEnv
    cfg = load(data=SCENES['CutRearrange-v1'])
    cfg.SHAPES[0]['n_particles'] = 700
    te = TaichiEnv(cfg, loss=False, max_env_steps=2)
    te.initialize()
    func = GradModel(te, softness=666.)
    x0 = torch.as_tensor(te.simulator.get_x(0), dtype=torch.float32, device=DEVICE)
    target = x0 + torch.tensor([0.0, -0.01, 0.02], device=DEVICE)

    def loss_fn(idx, x, c):
        d = ((x[:, :3] - target) ** 2).mean()
        knife = ((c[0, :3] - x[:, :3].mean(0)) ** 2).sum()
        return d + 0.1 * knife, {'dist': d.item(), 'knife': knife.item()}

    H, A = 2, te.primitives.action_dim
    init = np.zeros((H, A), np.float32)
    init[:, 1] = -0.3                                                   # knife goes down (solve_utils.py:166-168)
    before = te.get_state()
    st = solve(te, func, init, loss_fn, lr=0.05, max_iter=2, verbose=False, action_dims=(0, 1, 2), device=DEVICE)
    assert st['iter_id'] == 1 and len(st['optim_buffer']) == 2 and np.isfinite(st['best_loss'])
    assert (st['last_action'][:, 3:] == 0).all() and np.abs(st['last_action']).max() <= 1.0
    assert not np.allclose(st['last_action'][:, :3], init[:, :3])
    assert np.array_equal(te.get_state()['state'][0], before['state'][0])
    st2 = solve(te, func, None, loss_fn, max_iter=1, verbose=False, action_dims=(0, 1, 2), state=st, early_stop=5,
                compute_loss_in_end=True, device=DEVICE)
    assert len(st2['optim_buffer']) == 3 and st2['optim'] is st['optim'] and st2['best_loss'] <= st['best_loss']
    g = solve(te, func, np.full((H, A), 0.5, np.float32), lambda i, x, c: (x[:, :3] ** 2).mean(), max_iter=1, verbose=False,
              action_dims='gripper', device=DEVICE)['last_action']       # two 6-D tools tied in y / z (solve_func.py:110-119)
    assert np.allclose(g[:, [1, 2]], g[:, [7, 8]]) and (g[:, 3:6] == 0).all() and (g[:, 9:] == 0).all()


def test_multitask_env_over_a_cached_dataset(tmp_path):
    """plb/envs/multitask_env.py: init/state_<i>.xz + target/target_<i>.npy -> reset(init_v, target_v, contact_loss_mask) ->
    step -> reward / info; the dataset is written by envs.dataset.generate_synthetic in the reference's on-disk layout."""
    import torch
    from diffskill_b200.envs import MultitaskPlasticineEnv, make
    from diffskill_b200.envs.dataset import generate_synthetic, load_pair
    root = str(tmp_path / 'gathermove')
    gen = make('GatherMove-v1', max_env_steps=2)
    assert generate_synthetic(gen, root, 2, settle_steps=1) == [0, 1]
    st0, goal1 = load_pair(root, 0)[0], load_pair(root, 1)[1]
    del gen
    np.random.seed(0)
    env = MultitaskPlasticineEnv('GatherMove-v1', cached_state_path=root, device=DEVICE, max_env_steps=2)
    assert env.num_inits == 2 and env.num_targets == 2 and len(env.target_pcs) == 2 and env.action_dim == 13
    obs = env.reset(init_v=0, target_v=1, contact_loss_mask=[1., 0., 0.])
    te = env.taichi_env
    assert (env.init_v, env.target_v) == (0, 1) and np.array_equal(env.target_pc, goal1)
    assert np.allclose(te.simulator.get_x(0), st0['state'][0], atol=1e-7)           # the cached state, through fp32
    assert te.contact_loss_mask.tolist() == [1., 0., 0.] and te.init_emd > 0
    assert obs.ndim == 1 and np.isfinite(obs).all()
    a = np.zeros(13)
    a[1] = -2.0                                                                      # clipped to -1
    obs2, r, done, info = env.step(a)
    assert obs2.shape == obs.shape and np.isfinite(r) and done is False
    assert set(info) == {'info_emd', 'info_normalized_performance', 'info_contact_loss'}
    assert np.array_equal(env._recorded_actions[0], np.clip(a, -1, 1))
    s = env.get_state()
    env.step(np.zeros(13))
    env.set_state(s)
    assert np.array_equal(env.get_state()['state'][0], s['state'][0]) and len(env.get_primitive_state()) == 3
    env.reset(contact_loss_mask=0)                                                   # random pair, mask by tool index
    assert te.contact_loss_mask.tolist() == [1., 0., 0.] and env.init_v in (0, 1) and isinstance(te.tensor_target_x, torch.Tensor)


def test_mlp_policy_gradients_through_the_engine():
    """plb/engine/nn/mlp.py over the engine: TaichiEnv(nn=True).nn is the closed-loop policy (observation -> layers -> clamp ->
    action); its parameter gradients flow through GradModel's autograd nodes.  One env step: the flat get_grad() must equal
    (d action / d params)^T applied to the engine's own action gradient; two env steps: a directional finite difference
    through the whole closed loop."""
    import torch
    from diffskill_b200.config import load
    from diffskill_b200.envs.scenes import SCENES
    from diffskill_b200.sim import GradModel, MLP, TaichiEnv
    cfg = load(data=SCENES['LiftSpread-v1'])
    cfg.SHAPES[0]['radius'] = 0.02
    te = TaichiEnv(cfg, nn=True, loss=False, max_env_steps=3)
    te.initialize()
    assert isinstance(te.nn, MLP) and te.nn.dims[0] == te.nn.obs_num * 6 + 7 * len(te.primitives) and te.nn.dims[-1] == te.primitives.action_dim
    # dough in contact with the rolling pin and the lifter (tests/helpers.py), so that actions matter from the first step
    from helpers import small_dough, tool_start
    scene_, _, x0 = small_dough('LiftSpread-v1', te.n_particles, 0)
    te.simulator.reset(x0)
    for p, st in zip(te.primitives, tool_start('LiftSpread-v1', scene_)):
        p.set_state(0, list(st[:p.state_dim]))
    nn = MLP(te.simulator, te.primitives, (16,), activation='tanh', n_observed_particles=50, n_particles=te.n_particles, device=DEVICE, seed=1)
    flat = nn.get_params()
    nn.set_params(flat * 0.5)
    assert np.allclose(nn.get_params(), flat * 0.5) and len(flat) == sum(a * b + a for a, b in zip(nn.dims[1:], nn.dims[:-1]))
    func = GradModel(te, softness=666.)
    rng = np.random.RandomState(0)
    n = te.n_particles
    Wt = torch.tensor(rng.normal(size=(n, 6)), device=DEVICE, dtype=torch.float32)

    def loss_of(H):
        nn.zero_grad()
        return nn.rollout(func, H, lambda s, obs: (obs[0][:, :6] * Wt).sum() / n, device=DEVICE)

    # one step: chain rule by hand
    loss = loss_of(1)
    loss.backward()
    g = nn.get_grad()
    ga = torch.as_tensor(func.eng.get_action_grad(0)[0], device=DEVICE)
    obs0 = func.reset(device=DEVICE)
    nn.zero_grad()
    (nn.forward(obs0) * ga).sum().backward()
    g_manual = nn.get_grad()
    e1 = relerr(g, g_manual)
    # two steps: directional derivative of the closed loop
    loss2 = loss_of(2)
    loss2.backward()
    g2 = nn.get_grad()
    d = rng.normal(size=len(flat)) * (np.abs(g2) > 0)
    d /= np.linalg.norm(d)
    p0, fds = nn.get_params(), []
    for h in (1e-3, 2e-3, 4e-3):     # fp32 losses of O(1) differenced at 1e-5: average a few step sizes
        nn.set_params(p0 + h * d)
        lp = float(loss_of(2))
        nn.set_params(p0 - h * d)
        lm = float(loss_of(2))
        fds.append((lp - lm) / (2 * h))
    nn.set_params(p0)
    fd, an = float(np.mean(fds)), float(g2 @ d)
    print('MLP policy: 1-step chain err %.2e; 2-step directional derivative autograd %.4e vs FD %.4e' % (e1, an, fd))
    assert e1 < 1e-4 and np.abs(g).max() > 0
    assert an * fd > 0 and 1 / 3 < an / fd < 3, (an, fds)   # a consistency check at fp32 finite-difference noise, not a parity bar
