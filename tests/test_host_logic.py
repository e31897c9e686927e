"""Host-side logic that needs no GPU: config/scene derivation, synthetic dough, tool mirrors, sharding."""
import numpy as np
import pytest

from diffskill_b200.config import CfgNode, load
from diffskill_b200.envs.scenes import SCENES
from diffskill_b200.parallel import shard_envs
from diffskill_b200.scene import load_scene
from diffskill_b200.shapes import Shapes


def test_scene_constants_match_reference_derivation():
    s, _ = load_scene('LiftSpread-v1')
    assert (s.n_grid, s.substeps) == (64, 19) and s.dt == pytest.approx(1e-4)       # 2e-3 // 1e-4 == 19 (float floor)
    assert s.p_vol == (s.dx * 0.5) ** 2 and s.p_mass == s.p_vol                       # squared even in 3-D
    assert s.mu == pytest.approx(2173.913043) and s.lam == pytest.approx(931.677018)
    assert s.pairs == [(1, 2)] and s.action_dims == [0, 6, 12, 12]
    g, _ = load_scene('GatherMove-v1')
    assert g.pairs == [(0, 2), (1, 2)] and g.action_dim == 13 and g.tools[0].state_dim == 8
    c, _ = load_scene('CutRearrange-v1')
    assert (c.n_grid, c.substeps) == (80, 24) and c.dt == pytest.approx(8e-5) and c.lower_bound == 1.0
    assert c.pairs == [] and c.action_dims == [0, 3, 10] and c.tools[0].prot == (1.0, 0.0, 0.0, 0.58)
    assert c.yield_stress == 150. and c.gravity == (0., -10., 0.)


def test_string_tuples_are_decoded_like_yacs():
    cfg = load(data=dict(SIMULATOR=dict(gravity='(0, -20, 0)'), PRIMITIVES=[dict(shape='Box', size='(0.1, 0.2, 0.3)')]))
    assert cfg.SIMULATOR.gravity == (0, -20, 0)
    assert cfg.PRIMITIVES[0].size == (0.1, 0.2, 0.3)
    cfg.merge_from_list(['SIMULATOR.yield_stress', '75.'])
    assert cfg.SIMULATOR.yield_stress == 75.


def test_synthetic_dough_particle_counts():
    for name, n in (('LiftSpread-v1', 15707), ('GatherMove-v1', 2000), ('CutRearrange-v1', 5000)):
        _, cfg = load_scene(name)
        x, col = Shapes(cfg.SHAPES, seed=0).get()
        assert x.shape == (n, 3) and len(col) == n
        assert (x > 0).all() and (x < 1).all()
    a = Shapes(SCENES['GatherMove-v1']['SHAPES']).get()[0]
    b = Shapes(SCENES['GatherMove-v1']['SHAPES']).get()[0]
    assert np.array_equal(a, b)                                                       # scatter is seeded


def test_primitive_mirrors_unbound():
    from diffskill_b200.sim import Primitives
    cfg = load(data=SCENES['CutRearrange-v1'])
    P = Primitives(cfg.PRIMITIVES)
    assert len(P) == 2 and P.action_dim == 10 and P.state_dim == 15 and P.state_dims == [7, 8]
    assert P[1].init_state == (0.5, 0.10, 0.5, 0.707, 0.0, 0.707, 0.0, 0.18)
    assert P[0].get_state(0).shape == (7,) and P[1].get_state(0).shape == (8,)
    P[0].set_state(0, [0.1, 0.2, 0.3])
    assert np.allclose(P[0].get_state(0), [0.1, 0.2, 0.3, 1, 0, 0, 0])
    P.set_softness(10.)
    assert P.get_softness() == 10.
    P.initialize()
    assert P[1].init_points.shape == (100, 3) and P[0].init_points.shape == (100, 3)
    # inv_action: knife lifts first (primitives.py:790-791)
    a = P[0].inv_action(np.array([0.5, 0.1, 0.5, 1, 0, 0, 0.]), np.array([0.6, 0.3, 0.5, 1, 0, 0, 0.]))
    assert a[0] == 0. and a[1] == pytest.approx(8.0)
    assert P[0].inv_action(np.zeros(7), np.zeros(7)) is None
    with pytest.raises(NotImplementedError):
        Primitives([CfgNode(dict(shape='Spatula'))])            # not a PlasticineLab primitive
    Cs = Primitives([CfgNode(dict(shape='Chopsticks', action=dict(dim=7, scale=(0.02,) * 7)))])   # primitives.py:218-289
    assert Cs.state_dims == [8] and Cs[0].init_state[-1] == 0.06 and (Cs[0].spec.h, Cs[0].spec.r) == (0.06, 0.03)
    with pytest.raises(AssertionError):
        Primitives([CfgNode(dict(shape='Chopsticks'))])         # assert self.action_dim == 7 (primitives.py:228)
    # the legacy PlasticineLab tools with the defaults of their default_config()
    L = Primitives([CfgNode(dict(shape=s)) for s in ('Sphere', 'RollingPin', 'Cylinder', 'Torus')]
                   + [CfgNode(dict(shape='Gripper2', action=dict(dim=7, scale=(0.01,) * 7)))])
    assert L.state_dims == [7, 7, 7, 7, 8] and L.action_dim == 7
    assert (L[2].spec.h, L[2].spec.r) == (0.2, 0.1) and (L[3].spec.h, L[3].spec.r) == (0.2, 0.1)   # Torus: (tx, ty)
    assert L[4].init_state[-1] == 0.06 and L[4].spec.r == 0.015
    with pytest.raises(AssertionError):
        L[4].set_state(0, np.zeros(7))


def test_shard_envs_partitions_exactly():
    for total, world in ((64, 8), (64, 3), (256, 8), (5, 8)):
        blocks = [shard_envs(total, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b[1] - b[0] for b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_sinkhorn_stand_in_behaves_like_an_emd():
    """The plain-torch Sinkhorn used when geomloss is absent (taichi_env.py:23-26 stand-in)."""
    import torch
    from diffskill_b200.sim.taichi_env import sinkhorn_emd
    g = torch.Generator().manual_seed(0)
    x = torch.rand(120, 3, generator=g) * 0.1 + 0.4
    y = (x + torch.tensor([0.05, 0.0, 0.0])).clone()
    assert abs(float(sinkhorn_emd(x, x.clone()))) < 1e-4
    d = float(sinkhorn_emd(x, y))
    assert 0.04 < d < 0.06                                          # p=1: translation by 0.05 costs ~0.05
    assert abs(float(sinkhorn_emd(y, x)) - d) < 1e-3
    xr = x.clone().requires_grad_(True)
    sinkhorn_emd(xr, y).backward()
    gx = xr.grad
    assert torch.isfinite(gx).all()
    assert float(gx[:, 0].mean()) < 0                                # moving x towards +x lowers the loss


def test_dataset_format_round_trip(tmp_path):
    """init/state_i.xz (lzma pickle of TaichiEnv.get_state()) + target/target_i.npy, multitask_env.py:24-32,85-89."""
    from diffskill_b200.envs import dataset
    rng = np.random.RandomState(0)
    n = 50
    state = {'state': [rng.rand(n, 3), rng.rand(n, 3), rng.rand(n, 3, 3), rng.rand(n, 3, 3), rng.rand(7), rng.rand(8)],
             'softness': 666., 'is_copy': True}
    root = str(tmp_path / 'ds')
    dataset.save_pair(root, 3, state, rng.rand(n, 3))
    dataset.save_pair(root, 1, state, rng.rand(n, 3))
    assert dataset.list_pairs(root) == [1, 3]
    st, tgt = dataset.load_pair(root, 3)
    assert st['softness'] == 666. and st['is_copy'] is True and len(st['state']) == 6
    assert all(np.array_equal(a, b) for a, b in zip(st['state'], state['state']))
    assert tgt.shape == (n, 3) and tgt.dtype == np.float64
