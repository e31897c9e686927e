"""Host-side mirror of the reference's tool classes (plb/engine/primitive/primive_base.py,
primitives.py): same class names, constructor arguments, methods and error behaviour -- but no
arithmetic.  Signed distances, collision response, kinematics and their adjoints run in the CUDA
engine (csrc/tools.cuh, kernels_aux.cuh); these objects hold the configuration, slice actions and
forward state/parameter access to the C ABI once :class:`MPMSimulator` has bound them to an engine.
"""
import numpy as np
import yaml

from ..config import CfgNode, make_cls_config
from ..engine import PARAM_FRICTION, PARAM_LOWER, PARAM_SOFTNESS, PARAM_UPPER
from ..scene import TOOL_GRIPPER, ToolSpec, _primitive_defaults, tool_from_cfg
from .fields import ScalarField, ToolFrameField, ToolVectorField, ZeroOnFillGrad


def _quat_to_matrix(q):
    w, x, y, z = q
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def angvel(q1, q2):
    """Rotation vector taking q1 to q2 (primitive/utils.py:60-67), quaternions (w,x,y,z)."""
    a = np.array(q1, dtype=np.float64)
    b = np.array(q2, dtype=np.float64)
    ac = np.array([a[0], -a[1], -a[2], -a[3]])
    w = b[0] * ac[0] - b[1:] @ ac[1:]
    v = b[0] * ac[1:] + ac[0] * b[1:] + np.cross(b[1:], ac[1:])
    n = np.sqrt(w * w + v @ v)
    w, v = w / n, v / n
    ln = np.linalg.norm(v)
    return v * (2 * np.arctan2(ln, w))


class Primitive:
    """primive_base.py:13-329.  ``shape`` selects the SDF family inside the engine."""
    state_dim = 7
    shape = ''

    def __init__(self, cfg=None, dim=3, max_timesteps=1024, dtype='float32', **kwargs):
        self.cfg = make_cls_config(self, cfg, **kwargs)
        self.cfg.shape = self.shape
        self.dim = dim
        self.max_timesteps = max_timesteps
        self.dtype = dtype
        self.pos_dim, self.rotation_dim, self.angular_velocity_dim = dim, 4, 3
        self.action_dim = int(self.cfg.action.dim)
        self.num_rand_points = 100
        self.spec: ToolSpec = tool_from_cfg(self.cfg)
        self._engine, self._index, self._sim = None, None, None
        self._pending_state = np.array(self.spec.init_state, dtype=np.float64)
        self._friction, self._softness = float(self.cfg.friction), 666.
        self._limits = [list(self.spec.lower_bound), list(self.spec.upper_bound)]
        zero = self._zero_grads
        self.friction = ScalarField(lambda: self._friction, self._set_friction)
        self.softness = ScalarField(lambda: self._softness, self._set_softness)
        self.xyz_limit = ToolVectorField(self)
        self.position = ToolFrameField(self, slice(0, 3), zero)
        self.rotation = ToolFrameField(self, slice(3, 7), zero)
        for name in ('v', 'w', 'action_buffer', 'min_dist', 'dist_norm'):
            setattr(self, name, type('F', (), {'grad': ZeroOnFillGrad(zero)})())
        self.init_points = np.zeros((self.num_rand_points, 3))

    # ---- binding to the engine (done by MPMSimulator.__init__) -----------------------------------------
    def _bind(self, sim, index):
        self._sim, self._engine, self._index = sim, sim.engine, index
        for b in range(sim.n_envs):
            self._engine.set_tool_state(0, b, index, self._pending_state)
        self._set_friction(self._friction)
        self._set_softness(self._softness)

    def _zero_grads(self):
        if self._engine is not None:
            self._engine.zero_grad()

    def _set_friction(self, v):
        self._friction = float(v)
        if self._engine is not None:
            self._engine.set_tool_param(self._index, PARAM_FRICTION, self._friction)

    def _set_softness(self, v):
        self._softness = float(v)
        if self._engine is not None:
            self._engine.set_tool_param(self._index, PARAM_SOFTNESS, self._softness)

    def _get_limit(self, i):
        return self._limits[i]

    def _set_limit(self, i, v):
        self._limits[i] = [float(a) for a in v]
        if self._engine is not None:
            for d in range(3):
                self._engine.set_tool_param(self._index, (PARAM_LOWER if i == 0 else PARAM_UPPER) + d, self._limits[i][d])

    def _step_of(self, f):
        if self._sim is None:
            return 0
        return self._sim._frame_to_step(f)

    # ---- state ---------------------------------------------------------------------------------------------
    def get_state(self, f, env=0):
        if self._engine is None:
            return self._pending_state[:self.state_dim].copy()
        return self._engine.get_tool_state(self._step_of(f), env, self._index).astype(np.float64)[:self.state_dim]

    def set_state(self, f, state, env=None):
        ss = self.get_state(f)
        state = np.asarray(state, dtype=np.float64).reshape(-1)
        ss[:len(state)] = state                      # short states pad into the existing one, primive_base.py:188-191
        if self._engine is None:
            self._pending_state[:len(ss)] = ss
            return
        envs = range(self._sim.n_envs) if env is None else [env]
        for b in envs:
            self._engine.set_tool_state(self._step_of(f), b, self._index, ss)

    def get_state_tensor(self, f, device='cuda'):
        import torch
        return torch.tensor(self.get_state(f)[:7], dtype=torch.float32, device=device)

    @property
    def init_state(self):
        return tuple(self.cfg.init_pos) + tuple(self.cfg.init_rot)

    def initialize(self, cached_state_path=''):
        cfg = self.cfg
        self.set_state(0, self.init_state)
        self._set_limit(0, cfg.lower_bound)
        self._set_limit(1, cfg.upper_bound)
        self._set_friction(cfg.friction)
        self.generate_init_points(cached_state_path)

    def generate_init_points(self, cached_state_path=''):
        pass

    def update_cfg(self, cfg):
        self.cfg = make_cls_config(self, cfg)
        self.cfg.shape = self.shape
        self.spec = tool_from_cfg(self.cfg)
        self.set_state(0, self.init_state)

    def get_surface_points(self, f=0, state=None):
        if state is None:
            state = self.get_state(f)
        position, rotation = np.asarray(state[:3]), np.asarray(state[3:7])
        R = _quat_to_matrix(rotation)
        pts = self.init_points.copy()
        if len(state) == 8:
            pts[:50, 0] -= state[7] / 2
            pts[50:, 0] += state[7] / 2
        return pts @ R.T + position

    # ---- actions -------------------------------------------------------------------------------------------
    def set_action(self, s, n_substeps, action):
        """Single-tool variant (primive_base.py:270-274); Primitives.set_action is the batched path."""
        if self.action_dim > 0:
            self._sim._set_tool_action(self._index, s, np.asarray(action, dtype=np.float64).reshape(-1))

    def get_action_grad(self, s, n):
        if self.action_dim == 0:
            return None
        lo = self._sim.primitives.action_dims[self._index]
        g = self._engine.get_action_grads(s, n)[:, 0, lo:lo + self.action_dim]
        return g.astype(np.float64)

    def inv_action(self, curr_state, target_state, thr=1e-2):
        """primive_base.py:289-311."""
        curr_state, target_state = np.asarray(curr_state, float), np.asarray(target_state, float)
        thr = np.ones_like(curr_state) * thr
        if np.all(np.abs(curr_state - target_state) < thr):
            return None
        action = np.zeros(self.action_dim)
        d = (target_state[:3] - curr_state[:3]) * 40
        if np.any(np.abs(curr_state[:3] - target_state[:3]) > thr[:3]):
            action[:3] = d
        vel = angvel(curr_state[3:7], target_state[3:7])
        if vel[2] == 0.:
            vel[2] = -vel[1] * 0.3
            vel[1] = 0.
        vel = vel * 100.
        if np.any(np.abs(curr_state[3:7] - target_state[3:7]) > thr[3:7]) and self.action_dim >= 6:
            action[3:6] = vel
        return action

    @classmethod
    def default_config(cls):
        return _primitive_defaults(cls.shape)


class Sphere(Primitive):
    """primitives.py:23-41 (legacy PlasticineLab tool); the base class's zero init_points."""
    shape = 'Sphere'


class Capsule(Primitive):
    shape = 'Capsule'

    def generate_init_points(self, cached_state_path=''):   # primitives.py:68-76
        n = int(np.sqrt(self.num_rand_points))
        h, r = self.cfg.h, self.cfg.r
        l1, l2 = np.linspace(-h / 2, h / 2, n), np.linspace(0, 2 * np.pi, n)
        self.init_points = np.array([[r * np.cos(l2[k]), l1[l], r * np.sin(l2[k])] for k in range(n) for l in range(n)])


class RollingPin(Capsule):
    """primitives.py:101-117: rolls (dw), yaws about the world y (dth) and sinks (dy); 3-D action."""
    shape = 'RollingPin'


class RollingPinExt(Capsule):
    shape = 'RollingPinExt'

    def inv_action(self, curr_state, target_state, thr=1e-2):   # primitives.py:138-153
        curr_state, target_state = np.asarray(curr_state, float), np.asarray(target_state, float)
        thr = np.ones_like(curr_state) * thr
        if np.all(np.abs(curr_state - target_state)[:3] < thr[:3]):
            return None
        action, ret = np.zeros(self.action_dim), np.zeros(self.action_dim)
        if np.any(np.abs(curr_state[:3] - target_state[:3]) > thr[:3]):
            action[:3] = (target_state[:3] - curr_state[:3]) * 40
        ret[2] = action[1] * 50
        ret[0] = -action[0] * 0.2
        return ret


class Box(Primitive):
    shape = 'Box'

    def generate_init_points(self, cached_state_path=''):   # primitives.py:395-419
        n = int(np.sqrt(self.num_rand_points))
        size = self.cfg.size
        if 'cutrearrangespread' in cached_state_path:
            l1, l2 = np.linspace(-size[1], size[1], n), np.linspace(-size[2], size[2], n)
            pts = [[-size[0] / 2, l1[k], l2[l]] for k in range(n) for l in range(n)]
        else:
            l1, l2 = np.linspace(-size[0], size[0], n), np.linspace(-size[1], size[1], n)
            pts = [[l1[k], l2[l], -size[2] / 2] for k in range(n) for l in range(n)]
        self.init_points = np.array(pts)


class Gripper(Box):
    shape = 'Gripper'
    state_dim = 8

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        zero = self._zero_grads
        self.gap = ToolFrameField(self, slice(7, 8), zero)
        self.gap_vel = type('F', (), {'grad': ZeroOnFillGrad(zero)})()
        self.minimal_gap, self.maximal_gap = self.cfg.minimal_gap, self.cfg.maximal_gap

    @property
    def init_state(self):
        return tuple(self.cfg.init_pos) + tuple(self.cfg.init_rot) + (self.cfg.init_gap,)

    def set_state(self, f, state, env=None):
        assert len(state) == 8                                  # primitives.py:551
        super().set_state(f, state, env)

    def generate_init_points(self, cached_state_path=''):   # primitives.py:440-454
        n = int(np.sqrt(self.num_rand_points))
        size = self.cfg.size
        l1, l2 = np.linspace(-size[1], size[1], n), np.linspace(-size[2], size[2], n)
        pts = [[size[0], l1[l], l2[k]] for k in range(n) for l in range(n)]
        pts += [[-size[0], l1[l], l2[k]] for k in range(n) for l in range(n)]
        self.init_points = np.array(pts[::2])

    def inv_action(self, curr_state, target_state, thr=1e-2):   # primitives.py:555-561
        action = super().inv_action(curr_state, target_state, thr)
        if action is None:
            return None
        action[-1] = (-target_state[-1] + curr_state[-1]) * 50
        return action


class Gripper2(Capsule):
    """primitives.py:576-697: the Gripper's kinematics and two-jaw contact with capsule jaws (h, r)."""
    shape = 'Gripper2'
    state_dim = 8

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        zero = self._zero_grads
        self.gap = ToolFrameField(self, slice(7, 8), zero)
        self.gap_vel = type('F', (), {'grad': ZeroOnFillGrad(zero)})()
        self.minimal_gap, self.maximal_gap = self.cfg.minimal_gap, self.cfg.maximal_gap

    @property
    def init_state(self):
        return tuple(self.cfg.init_pos) + tuple(self.cfg.init_rot) + (self.cfg.init_gap,)

    def set_state(self, f, state, env=None):
        assert len(state) == 8                                  # primitives.py:679
        super().set_state(f, state, env)


class Chopsticks(Capsule):
    """primitives.py:218-300: two capsules (h, r) at -+gap/2 along local x inside ONE tool frame -- sdf = min, the nearer
    stick's normal, the base class's single contact (collider velocity from the tool pose alone) -- with the gripper's
    kinematics (right-multiplied rotation, gap' = max(gap - gap_vel, minimal_gap)) and 7-D action."""
    shape = 'Chopsticks'
    state_dim = 8
    dist_cols = 1                                               # one contact-distance column (sdf = min of the sticks)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        zero = self._zero_grads
        self.gap = ToolFrameField(self, slice(7, 8), zero)
        self.gap_vel = type('F', (), {'grad': ZeroOnFillGrad(zero)})()
        self.minimal_gap = self.cfg.minimal_gap
        assert self.action_dim == 7                             # primitives.py:228: 3 linear, 3 angle, 1 for grasp

    @property
    def init_state(self):
        return tuple(self.cfg.init_pos) + tuple(self.cfg.init_rot) + (self.cfg.init_gap,)

    def set_state(self, f, state, env=None):
        assert len(state) == 8                                  # primitives.py:274
        super().set_state(f, state, env)


class Cylinder(Primitive):
    """primitives.py:302-336 (cfg.h = radial, cfg.r = axial half extent); the base class's zero init_points."""
    shape = 'Cylinder'


class Torus(Primitive):
    """primitives.py:337-365 (cfg.tx major, cfg.ty minor radius)."""
    shape = 'Torus'


class Knife(Primitive):
    shape = 'Knife'

    def generate_init_points(self, cached_state_path=''):   # primitives.py:752-767
        h, size = self.cfg.h, self.cfg.size
        n = int(np.sqrt(self.num_rand_points))
        l1, l2 = np.linspace(0, size[0], n), np.linspace(-size[2], size[2], n)
        p1 = [[-l1[l], l1[l] * np.sqrt(3) - h[0], l2[k]] for k in range(n) for l in range(n)][::2]
        p2 = [[l1[l], l1[l] * np.sqrt(3) - h[0], l2[k]] for k in range(n) for l in range(n)][::2]
        self.init_points = np.array(p1 + p2)

    def inv_action(self, curr_state, target_state, thr=1e-2):   # primitives.py:780-794
        curr_state, target_state = np.asarray(curr_state, float), np.asarray(target_state, float)
        thr = np.ones_like(curr_state) * thr
        if np.all(np.abs(curr_state - target_state) < thr):
            return None
        action = np.zeros(self.action_dim)
        d = (target_state[:3] - curr_state[:3]) * np.array([10., 40., 10.])
        if abs(target_state[1] - curr_state[1]) > 0.01:
            d[0] = 0.
        if np.any(np.abs(curr_state[:3] - target_state[:3]) > thr[:3]):
            action[:3] = d
        return action


_SHAPES = {c.shape: c for c in (Sphere, Capsule, RollingPin, RollingPinExt, Box, Gripper, Gripper2, Cylinder, Torus, Knife, Chopsticks)}


class Primitives:
    """primitives.py:820-897."""

    def __init__(self, cfgs, max_timesteps=1024):
        self.primitives = []
        self.action_dims = [0]
        for i in cfgs:
            cfg = i if isinstance(i, CfgNode) else CfgNode(yaml.safe_load(yaml.safe_dump(dict(i))))
            if cfg.shape not in _SHAPES:
                raise NotImplementedError(f"primitive {cfg.shape!r} is not built: every tool of primitives.py except "
                                          "Chopsticks is (SURVEY.md section 8f row 4)")
            p = _SHAPES[cfg.shape](cfg=cfg, max_timesteps=max_timesteps)
            self.primitives.append(p)
            self.action_dims.append(self.action_dims[-1] + p.action_dim)
        self.n = len(self.primitives)
        self._sim = None

    def update_cfgs(self, cfgs):
        for p, c in zip(self.primitives, cfgs):
            p.update_cfg(c if isinstance(c, CfgNode) else CfgNode(dict(c)))

    @property
    def action_dim(self):
        return self.action_dims[-1]

    @property
    def state_dim(self):
        return sum(i.state_dim for i in self.primitives)

    @property
    def state_dims(self):
        return [i.state_dim for i in self.primitives]

    def set_action(self, s, n_substeps, action):
        action = np.asarray(action).reshape(-1).clip(-1, 1)
        assert len(action) == self.action_dims[-1]              # primitives.py:865
        self._sim._set_action(s, action)

    def get_grad(self, n):
        g = self._sim.engine.get_action_grads(0, n)[:, 0, :]
        return g.astype(np.float64)

    def set_softness(self, softness=666.):
        for i in self.primitives:
            i.softness[None] = softness

    def get_softness(self):
        return self.primitives[0].softness[None]

    def __getitem__(self, item):
        if isinstance(item, tuple):
            item = item[0]
        return self.primitives[item]

    def __len__(self):
        return len(self.primitives)

    def __iter__(self):
        return iter(self.primitives)

    def initialize(self, cached_state_path=''):
        for i in self.primitives:
            i.initialize(cached_state_path)
