"""Dough generators (plb/engine/shapes/shape_maker.py) and the start/goal generators of the batched configs
(core/diffskill/gen_init_target/gathermove_generator_V2.py, cutrearrange_generator_0528.py) -- the data either side of the
hot path.  CPU only."""
import numpy as np

from diffskill_b200.envs import generators as gen
from diffskill_b200.scene import load_scene
from diffskill_b200.shapes import (Shapes, get_n_particles, make_capsule, make_cylinder, make_multibox, make_multisphere,
                                   quat_to_mat, rotate_about_centroid)


def test_particle_count_rule_and_lift_spread_dough():
    assert get_n_particles((0.05 ** 3) * 4 * np.pi / 3) == 15707          # shape_maker.py:56 on LiftSpread's sphere
    scene, cfg = load_scene('LiftSpread-v1')
    x, c = Shapes(cfg.SHAPES).get()
    assert x.shape == (15707, 3) and c.shape == (15707,)
    r = np.linalg.norm(x - np.array(cfg.SHAPES[0]['init_pos']), axis=1)
    assert r.max() <= cfg.SHAPES[0]['radius'] + 1e-9


def test_cylinder_and_capsule_geometry():
    rng = np.random.RandomState(0)
    p = make_cylinder((0.5, 0.2, 0.5), 0.03, 0.1, 4000, rng)
    d = p - np.array([0.5, 0.2, 0.5])
    assert np.hypot(d[:, 0], d[:, 1]).max() <= 0.03 + 1e-9 and np.abs(d[:, 2]).max() <= 0.05 + 1e-9   # axis = z
    q = make_capsule((0.5, 0.2, 0.5), 0.03, 0.1, 6000, rng)
    d = q - np.array([0.5, 0.2, 0.5])
    assert len(q) == 6000 and np.abs(d[:, 2]).max() <= 0.05 + 0.03 + 1e-9
    # points beyond the cylinder's ends lie on the two half balls
    cap = np.abs(d[:, 2]) > 0.05
    assert (np.linalg.norm(d[cap] - np.sign(d[cap, 2:3]) * np.array([0, 0, 0.05]), axis=1) <= 0.03 + 1e-9).all()


def test_rotation_about_the_centroid():
    R = quat_to_mat((0.70710678, 0.70710678, 0., 0.))                     # 90 degrees about x
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and np.allclose(R @ np.array([0, 0, 1.]), [0, -1, 0], atol=1e-6)
    p = make_cylinder((0.5, 0.2, 0.5), 0.02, 0.2, 2000, np.random.RandomState(1))
    q = rotate_about_centroid(p, (0.70710678, 0.70710678, 0., 0.))
    assert np.allclose(q.mean(0), p.mean(0), atol=1e-9)
    assert np.ptp(q[:, 1]) > 0.19 and np.ptp(q[:, 2]) < 0.041            # the axis now points along y


def test_multi_shapes_share_one_particle_budget():
    boxes = make_multibox([(0.3, 0.1, 0.5), (0.6, 0.1, 0.5)], [(0.04, 0.04, 0.04), (0.08, 0.04, 0.04)])
    assert abs(len(boxes[1]) - 2 * len(boxes[0])) <= 2                    # proportional to the volumes
    assert sum(map(len, boxes)) <= 30000
    balls = make_multisphere([(0.3, 0.1, 0.5), (0.6, 0.1, 0.5)], [0.02, 0.04])
    assert abs(len(balls[1]) - 8 * len(balls[0])) <= 8
    x, c = Shapes([dict(shape='multisphere', all_pos=[(0.3, 0.1, 0.5), (0.6, 0.1, 0.5)], all_r=[0.02, 0.04]),
                   dict(shape='box', init_pos='(0.5, 0.1, 0.5)', width=0.03, color=7)]).get()
    assert len(x) == len(c) and set(np.unique(c)) >= {7}


def test_gathermove_starts_and_goals():
    scene, cfg = load_scene('GatherMove-v1')
    a, b = gen.gathermove_start(cfg, 3), gen.gathermove_start(cfg, 3)
    assert a.shape == (2000, 3) and np.array_equal(a, b)                  # seed = index (gathermove_generator_V2.py:13-17)
    assert not np.array_equal(a, gen.gathermove_start(cfg, 4))
    lo, hi = np.array(cfg.SHAPES[0]['pos_min']), np.array(cfg.SHAPES[0]['pos_max'])
    assert (a[:, [0, 2]] > lo[[0, 2]] - 0.03).all() and (a[:, [0, 2]] < hi[[0, 2]] + 0.03).all()
    g = gen.gathermove_goal(3, 2000)
    xs, rs = np.linspace(0.36, 0.4, 100), np.linspace(0.04, 0.07, 10)
    r = rs[3 * 11117771 % 12837119 % 100 // 10]
    centre = np.array([xs[3], r + 0.08, 0.5])
    assert g.shape == (2000, 3) and np.linalg.norm(g - centre, axis=1).max() <= r + 1e-6


def test_cutrearrange_pairs():
    rng = np.random.RandomState(0)
    x, cut, flag, (ta, tb) = gen.cutrearrange_pair(rng)
    assert x.shape == (5000, 3) and cut.shape == (5000, 3) and flag.dtype == bool
    w = np.ptp(x, axis=0)
    assert 0.19 < w[0] < 0.241 and 0.07 < w[1] < 0.081 and 0.03 < w[2] < 0.081      # width (U[.2,.24], .08, U[.04,.08])
    moved = np.abs(cut[:, 0] - x[:, 0]) > 1e-9
    assert moved.any() and (moved == flag).all() or (moved == ~flag).all()           # exactly one half slides, by 0.1-0.13
    assert 0.0999 < np.abs(cut[:, 0] - x[:, 0]).max() < 0.1301
    assert np.array_equal(cut[:, 1:], x[:, 1:])
    assert np.linalg.norm(np.array(ta) - np.array(tb)) >= 0.3                        # two far-apart targets
    out, mean = gen.move_cluster(cut, flag, *ta)
    assert np.allclose(out[flag].mean(0)[[0, 2]], ta, atol=1e-6) and np.array_equal(out[~flag], cut[~flag])
    a = gen.knife_init_actions(42)
    assert a.shape == (42, 10) and (a[:20, 1] == -0.3).all() and a[20:].sum() == 0 and a[:, [0, 2]].sum() == 0
