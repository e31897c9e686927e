"""Synthetic start/goal pairs for the batched configs.

The Google-Drive dataset is unavailable offline; the reference generates it with the two scripts restated here:

* ``core/diffskill/gen_init_target/gathermove_generator_V2.py:13-32`` -- GatherMove: env i's dough is the `scatter` shape
  with seed i, left to settle for 10 zero-action env steps; goals are sphere blobs (`case2`, :19-26).
* ``core/diffskill/gen_init_target/cutrearrange_generator_0528.py:7-81`` -- CutRearrange: a 5 000-particle box of random
  width, cut at a random x and slid apart; two far-apart targets for the two halves.

Both draw from numpy's global RNG in the reference; here every draw comes from an explicit ``RandomState``.
"""
import numpy as np

from ..shapes import make_scatter, make_sphere


# ---- GatherMove ---------------------------------------------------------------------------------------------------
GATHERMOVE_N = 100   # gathermove_generator_V2.py:9-11


def gathermove_start(cfg, i):
    """Particles of start i before settling: SHAPES[0] with seed = i (:13-17)."""
    s = dict(cfg.SHAPES[0])
    return make_scatter(s['pos_min'], s['pos_max'], i).astype(np.float32)


def gathermove_goal(i, n_particles, rng=None):
    """Goal i: a sphere blob (`case2`, :19-26) of radius rs[...] sitting at (xs[i], r + 0.08, 0.5).  The reference samples
    it at the volume-derived particle count; the benchmark's per-particle L2 stand-in loss needs the dough's own count, so
    `n_particles` is explicit."""
    xs = np.linspace(0.36, 0.4, GATHERMOVE_N)
    rs = np.linspace(0.04, 0.07, GATHERMOVE_N // 10)
    i = i % GATHERMOVE_N
    r = rs[i * 11117771 % 12837119 % GATHERMOVE_N // 10]
    return make_sphere((xs[i], r + 0.08, 0.5), r, n_particles, rng or np.random.RandomState(1000 + i)).astype(np.float32)


GATHERMOVE_SETTLE_STEPS = 10   # :30-32 "Wait for dough to drop"


def settle(eng, n_steps=GATHERMOVE_SETTLE_STEPS):
    """Zero-action env steps in copy mode (frame S -> 0) on every env of a batched engine; checkpoint 0 then holds the
    settled state (x, v, F, C), which is what the reference saves as the init."""
    eng.set_action(0, np.zeros((eng.B, eng.A), np.float32))
    for _ in range(n_steps):
        eng.forward_step(0, 0, 0)


# ---- CutRearrange -------------------------------------------------------------------------------------------------
def cutrearrange_pair(rng, n_particles=5000):
    """One start/goal pair as cutrearrange_generator_0528.py:40-81: returns (start x[N,3], cut x[N,3], flag[N], targets)
    where `cut` is the goal of the cutting skill (halves slid apart) and targets are the two far-apart placements."""
    def rand(a, b):
        return rng.random_sample() * (b - a) + a

    width = np.array([rand(0.2, 0.24), 0.08, rand(0.04, 0.08)])
    x = (rng.random_sample((n_particles, 3)) * 2 - 1) * (0.5 * width) + np.array([0.5, 0.0669, 0.5])
    da = db = 0.
    if rng.randint(2):
        da = rand(0.1, 0.13)
    else:
        db = rand(0.1, 0.13)
    cut_loc = rand(0.48, 0.52)
    flag = x[:, 0] <= cut_loc
    cut = x.copy()
    cut[flag, 0] -= da
    cut[~flag, 0] += db

    def sample_target():
        y = rand(0.3, 0.35) * (rng.randint(2) * 2 - 1) + 0.5
        return (rand(0.3, 0.7), y)

    while True:   # two targets which are far away
        a, b = sample_target(), sample_target()
        if np.linalg.norm(np.array(a) - np.array(b)) >= 0.3:
            break
    return x.astype(np.float32), cut.astype(np.float32), flag, (a, b)


def move_cluster(x, flag, dx, dz, dy=0.):
    """:16-22: translate the flagged half so that its mean lands on (dx, mean_y + dy, dz)."""
    out = x.copy()
    mean = out[flag].mean(axis=0)
    out[flag] += np.array([dx, mean[1] + dy, dz]) - mean
    return out, mean


def knife_init_actions(horizon, action_dim=10, clever_init=True):
    """plb/cut/solve_utils.py:163-168: the cutting skill starts from a 20-step downward push of the knife."""
    init = np.zeros((horizon, action_dim), np.float32)
    if clever_init:
        init[:20, 1] = -0.3
    return init
