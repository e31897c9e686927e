"""Shared helpers for the test-suite (synthetic dough near the tools, losses, error metrics)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from diffskill_b200.scene import load_scene  # noqa: E402
from diffskill_b200.shapes import Shapes  # noqa: E402


# the three DiffSkill envs, then legacy PlasticineLab scenes that exercise the remaining tools:
# Move-v1 Sphere, Rollingpin-v1 RollingPin, Torus-v1 Torus, Rope-v1 Sphere + Cylinder, Gripper2-synthetic Gripper2
ENVS = ['LiftSpread-v1', 'GatherMove-v1', 'CutRearrange-v1', 'Move-v1', 'Rollingpin-v1', 'Torus-v1', 'Rope-v1',
        'Gripper2-synthetic', 'Chopsticks-v1']


def small_dough(name, n, seed=0):
    """n particles of the env's synthetic dough, squeezed next to the tools so contacts are active."""
    scene, cfg = load_scene(name)
    rng = np.random.RandomState(seed)
    if name == 'LiftSpread-v1':
        # blob on the lifter plate, touching the rolling pin's influence region
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.04, 0.035, 0.04]) + np.array([0.62, 0.10, 0.5])
    elif name == 'GatherMove-v1':
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.05, 0.012, 0.05]) + np.array([0.70, 0.05, 0.5])
    elif name == 'CutRearrange-v1':
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.06, 0.03, 0.03]) + np.array([0.5, 0.06, 0.5])
    elif name == 'Move-v1':
        # slab between the two sphere manipulators (x = 0.576 and 0.776, radius 0.03), 5 mm into each of them
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.075, 0.03, 0.03]) + np.array([0.6757143, 0.5619162, 0.7515980])
    elif name == 'Rollingpin-v1':
        # slab under the pin (capsule r = 0.03 at y = 0.123, axis along world z): top face 7 mm inside it
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.04, 0.015, 0.06]) + np.array([0.5, 0.085, 0.5])
    elif name == 'Torus-v1':
        # slab under the ring (major 0.05, minor 0.03, axis = world y), which tool_start lowers to y = 0.12
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.09, 0.02, 0.09]) + np.array([0.5, 0.08, 0.5])
    elif name == 'Rope-v1':
        # piece of rope pressed 9 mm into the side of the pillar (Cylinder radius 0.1 at z = 0.499), the two spheres
        # (moved by tool_start) touching it from the other side
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.05, 0.03, 0.02]) + np.array([0.392, 0.04, 0.61])
    elif name == 'Chopsticks-v1':
        # a piece of the rope between the two sticks (capsules spanning y in [pos.y - h, pos.y] at x = pos.x -+ gap/2, r = 0.02),
        # 4 mm into each of them
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.024, 0.02, 0.06]) + np.array([0.5, 0.05, 0.5])
    elif name == 'Gripper2-synthetic':
        # slab between the two capsule jaws (axis = world y, separated along world z), 5 mm into each
        x = rng.uniform(-1, 1, (n, 3)) * np.array([0.04, 0.03, 0.035]) + np.array([0.5, 0.06, 0.5])
    else:
        x = Shapes(cfg.SHAPES, seed=seed).get()[0][:n]
    return scene, cfg, x


def perturbed_state(x, seed=1, vel=0.3, strain=0.02):
    """A generic (non-degenerate) particle state: small random v, C, and F = I + noise."""
    rng = np.random.RandomState(seed)
    n = len(x)
    v = rng.normal(size=(n, 3)) * vel
    C = rng.normal(size=(n, 3, 3)) * 5.0
    F = np.eye(3)[None] + rng.normal(size=(n, 3, 3)) * strain
    return v, F, C


def tool_start(name, scene):
    """Tool states at frame 0 moved so that every tool touches the small dough."""
    st = [np.array(t.init_state, dtype=np.float64) for t in scene.tools]
    if name == 'LiftSpread-v1':
        st[0][:3] = (0.60, 0.16, 0.5)     # rolling pin (y clamped to >= 0.16 by its lower_bound) 5 mm into the blob
        st[1][:3] = (0.62, 0.05, 0.5)     # lifter plate 5 mm into its underside
    elif name == 'GatherMove-v1':
        st[0][7] = 0.12                   # gripper jaws close to the dough
    elif name == 'CutRearrange-v1':
        st[0][:3] = (0.5, 0.24, 0.5)      # knife tip inside the slab
        st[1][:3] = (0.5, 0.09, 0.5)
        st[1][7] = 0.10
    elif name == 'Torus-v1':
        st[0][:3] = (0.5, 0.12, 0.5)
    elif name == 'Rope-v1':
        st[0][:3] = (0.36, 0.04, 0.655)
        st[1][:3] = (0.42, 0.04, 0.655)
    elif name == 'Chopsticks-v1':
        st[0][:3] = (0.5, 0.17, 0.5)      # sticks reach down to y = -0.03 .. 0.17 around the rope piece
        st[0][7] = 0.08
    elif name == 'Gripper2-synthetic':
        st[0][7] = 0.10
    return st


def relerr(a, b, floor=1e-30):
    """normwise relative error max|a-b| / max(max|b|, floor)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))
