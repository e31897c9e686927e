"""Scene description shared by the CUDA engine and (in tests) the oracle.

Restates the derived constants of ``MPMSimulator.__init__``
(plb/engine/mpm_simulator.py:8-67) and the per-class defaults of the tools
(plb/engine/primitive/primive_base.py:313-329, primitives.py:36-40, 86-91,
421-425, 563-573, 809-815).  Everything here is host-side double arithmetic,
exactly as the reference evaluates it in Python before Taichi rounds to fp32.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from .config import CfgNode, load

(TOOL_CAPSULE, TOOL_ROLLINGPIN_EXT, TOOL_BOX, TOOL_GRIPPER, TOOL_KNIFE, TOOL_SPHERE, TOOL_ROLLINGPIN, TOOL_GRIPPER2,
 TOOL_CYLINDER, TOOL_TORUS, TOOL_CHOPSTICKS) = range(11)
TOOL_TYPE = {'Capsule': TOOL_CAPSULE, 'RollingPinExt': TOOL_ROLLINGPIN_EXT, 'Box': TOOL_BOX,
             'Gripper': TOOL_GRIPPER, 'Knife': TOOL_KNIFE, 'Sphere': TOOL_SPHERE, 'RollingPin': TOOL_ROLLINGPIN,
             'Gripper2': TOOL_GRIPPER2, 'Cylinder': TOOL_CYLINDER, 'Torus': TOOL_TORUS, 'Chopsticks': TOOL_CHOPSTICKS}
GRIPPER_LIKE = (TOOL_GRIPPER, TOOL_GRIPPER2)      # two jaws applied one after the other (primitives.py:428, :576)
HAS_GAP = GRIPPER_LIKE + (TOOL_CHOPSTICKS,)       # gap state, 8-float state, 7-D action (+ Chopsticks, primitives.py:218)
NUM_COLLISION_POINTS = 600  # mpm_simulator.py:59


def _primitive_defaults(shape):
    d = dict(shape=shape, init_pos=(0.3, 0.3, 0.3), init_rot=(1., 0., 0., 0.), color=(0.3, 0.3, 0.3),
             lower_bound=(0., 0., 0.), upper_bound=(1., 1., 1.), friction=0.9, collision_group=[0., 0., 0.],
             action=dict(dim=0, scale=()))
    if shape in ('Capsule', 'RollingPinExt', 'RollingPin'):
        d.update(h=0.06, r=0.03)
    elif shape == 'Sphere':
        d.update(radius=1.)
    elif shape == 'Box':
        d.update(size=(0.1, 0.1, 0.1))
    elif shape == 'Gripper':
        d.update(size=(0.03, 0.06, 0.03), minimal_gap=0.06, maximal_gap=1., init_gap=0.06, round=0)
    elif shape == 'Gripper2':                                  # primitives.py:685-696
        d.update(h=0.06, r=0.015, minimal_gap=0.06, maximal_gap=1., init_gap=0.06, round=0)
    elif shape == 'Cylinder':                                  # primitives.py:331-336
        d.update(h=0.2, r=0.1)
    elif shape == 'Torus':                                     # primitives.py:360-365
        d.update(tx=0.2, ty=0.1)
    elif shape == 'Knife':
        d.update(h=(0.1, 0.1), size=(0.1, 0.1, 0.1), prot=(1.0, 0.0, 0.0, 0.0))
    elif shape == 'Chopsticks':                                # primitives.py:282-289
        d.update(h=0.06, r=0.03, minimal_gap=0.06, init_gap=0.06)
    else:
        raise NotImplementedError(f"tool shape {shape!r} is not a PlasticineLab primitive")
    return CfgNode(d)


@dataclass
class ToolSpec:
    shape: str
    type_id: int
    cfg: CfgNode
    action_dim: int
    action_scale: Tuple[float, ...]
    friction: float
    lower_bound: Tuple[float, float, float]
    upper_bound: Tuple[float, float, float]
    size: Tuple[float, float, float] = (0., 0., 0.)
    h: float = 0.
    r: float = 0.
    radius: float = 0.
    prism_h: Tuple[float, float] = (0., 0.)
    prot: Tuple[float, float, float, float] = (1., 0., 0., 0.)
    minimal_gap: float = 0.
    maximal_gap: float = 0.
    init_state: Tuple[float, ...] = ()
    collision_group: Tuple[float, ...] = ()

    @property
    def state_dim(self):
        return 8 if self.type_id in HAS_GAP else 7


def tool_from_cfg(c) -> ToolSpec:
    cfg = _primitive_defaults(c['shape'])
    cfg.merge_from_other_cfg(CfgNode(c) if not isinstance(c, CfgNode) else c)
    shape = cfg.shape
    t = TOOL_TYPE[shape]
    scale = tuple(float(v) for v in cfg.action.scale)
    adim = int(cfg.action.dim)
    assert len(scale) >= adim, f"{shape}: action.scale shorter than action.dim"
    spec = ToolSpec(shape=shape, type_id=t, cfg=cfg, action_dim=adim, action_scale=scale,
                    friction=float(cfg.friction), lower_bound=tuple(map(float, cfg.lower_bound)),
                    upper_bound=tuple(map(float, cfg.upper_bound)),
                    collision_group=tuple(float(v) for v in cfg.collision_group))
    init = tuple(map(float, cfg.init_pos)) + tuple(map(float, cfg.init_rot))
    if t in (TOOL_CAPSULE, TOOL_ROLLINGPIN_EXT, TOOL_ROLLINGPIN, TOOL_CYLINDER):
        spec.h, spec.r = float(cfg.h), float(cfg.r)
    elif t == TOOL_TORUS:                                      # major / minor radius travel in (h, r)
        spec.h, spec.r = float(cfg.tx), float(cfg.ty)
    elif t == TOOL_GRIPPER2:
        spec.h, spec.r = float(cfg.h), float(cfg.r)
        spec.minimal_gap, spec.maximal_gap = float(cfg.minimal_gap), float(cfg.maximal_gap)
        init = init + (float(cfg.init_gap),)
        assert adim == 7, "Gripper2 needs a 7-D action (primitives.py:594-601)"
    elif t == TOOL_CHOPSTICKS:                                 # primitives.py:218-289: no maximal gap
        spec.h, spec.r = float(cfg.h), float(cfg.r)
        spec.minimal_gap, spec.maximal_gap = float(cfg.minimal_gap), 1e30
        init = init + (float(cfg.init_gap),)
        assert adim == 7, "Chopsticks needs a 7-D action (primitives.py:228)"
    elif t == TOOL_SPHERE:
        spec.radius = float(cfg.radius)
    elif t == TOOL_BOX:
        spec.size = tuple(map(float, cfg.size))
    elif t == TOOL_GRIPPER:
        spec.size = tuple(map(float, cfg.size))
        spec.minimal_gap, spec.maximal_gap = float(cfg.minimal_gap), float(cfg.maximal_gap)
        init = init + (float(cfg.init_gap),)
        assert adim == 7, "Gripper needs a 7-D action (primitives.py:462-469)"
    elif t == TOOL_KNIFE:
        spec.size = tuple(map(float, cfg.size))
        spec.prism_h = tuple(map(float, cfg.h))
        spec.prot = tuple(map(float, cfg.prot))
    if len(init) == 7:
        init = init + (0.,)
    spec.init_state = init
    return spec


@dataclass
class SceneSpec:
    n_grid: int
    dx: float
    inv_dx: float
    dt: float
    p_vol: float
    p_mass: float
    substeps: int
    E: float
    nu: float
    mu: float
    lam: float
    yield_stress: float
    gravity: Tuple[float, float, float]
    ground_friction: float
    lower_bound: float
    particle_capacity: int
    max_steps: int
    tools: List[ToolSpec] = field(default_factory=list)
    pairs: List[Tuple[int, int]] = field(default_factory=list)
    shapes: list = field(default_factory=list)
    env_name: str = ''

    @property
    def action_dims(self):
        out = [0]
        for t in self.tools:
            out.append(out[-1] + t.action_dim)
        return out

    @property
    def action_dim(self):
        return self.action_dims[-1]

    def rand_num(self):
        """mpm_simulator.py:89-97: RandomState(42).uniform(-1.5, 1.5, (pairs, 600, 3))."""
        if not self.pairs:
            return np.zeros((0, NUM_COLLISION_POINTS, 3))
        return np.random.RandomState(42).uniform(-1.5, 1.5, (len(self.pairs), NUM_COLLISION_POINTS, 3))


def scene_from_cfg(cfg) -> SceneSpec:
    sim = cfg.SIMULATOR
    assert int(sim.dim) == 3, "only the 3-D path is on the DiffSkill hot path"
    assert sim.dtype == 'float32', "hot path is fp32 (SIMULATOR.dtype, default_config.py:15)"
    quality = sim.quality * sim.quality_multiplier * 0.5      # mpm_simulator.py:19-21
    n_grid = int(128 * quality)                                # :24
    dx, inv_dx = 1 / n_grid, float(n_grid)                     # :26
    dt = 0.5e-4 / quality                                      # :27
    p_vol = (dx * 0.5) ** 2                                    # :28 (squared also in 3-D)
    p_mass = p_vol * 1
    E, nu = float(sim.E), float(sim.nu)
    mu, lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))   # :33
    substeps = int(2e-3 // dt)                                 # :39 -> 19 (n=64), 24 (n=80)
    tools = [tool_from_cfg(c) for c in cfg.PRIMITIVES]
    pairs = []                                                 # :69-80
    for i, ti in enumerate(tools):
        for j in range(len(tools)):
            if j < len(ti.collision_group) and ti.collision_group[j] > 0:
                assert ti.type_id != TOOL_CHOPSTICKS, "Chopsticks as the moving tool of a tool-tool collision pair is not built"
                assert tools[j].type_id in (TOOL_BOX, TOOL_GRIPPER, TOOL_KNIFE), \
                    "tool-tool collision needs a box-like obstacle (mpm_simulator.py:291)"
                pairs.append((i, j))
    return SceneSpec(n_grid=n_grid, dx=dx, inv_dx=inv_dx, dt=dt, p_vol=p_vol, p_mass=p_mass, substeps=substeps,
                     E=E, nu=nu, mu=mu, lam=lam, yield_stress=float(sim.yield_stress),
                     gravity=tuple(float(g) for g in sim.gravity), ground_friction=float(sim.ground_friction),
                     lower_bound=float(sim.lower_bound), particle_capacity=int(sim.n_particles),
                     max_steps=int(sim.max_steps), tools=tools, pairs=pairs, shapes=list(cfg.SHAPES),
                     env_name=cfg.ENV.env_name if 'env_name' in cfg.ENV else '')


def load_scene(name_or_path, opts=None) -> Tuple[SceneSpec, CfgNode]:
    """Scene by registered env name ('LiftSpread-v1', ...) or by YAML path."""
    from .envs.scenes import SCENES
    if name_or_path in SCENES:
        cfg = load(data=SCENES[name_or_path], opts=opts)
    else:
        cfg = load(path=name_or_path, opts=opts)
    return scene_from_cfg(cfg), cfg
