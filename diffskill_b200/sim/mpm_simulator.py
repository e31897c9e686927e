"""``MPMSimulator`` with the reference's interface (plb/engine/mpm_simulator.py:6-488) over the CUDA engine.

Same constructor, attributes and method names; every method forwards to the C ABI
(include/diffskill_mpm.h).  Differences that follow from the B200 design are confined to storage:

* the reference tapes every substep frame (``max_steps = 1024``); the engine keeps checkpoints at
  env-step boundaries plus a ring of step slots, so ``get_state(f)`` / ``get_x(f)`` accept frames that
  are step boundaries (``f % substeps == 0``) or still resident in the ring;
* ``substep`` / ``substep_grad`` must follow the reference's own access pattern (ascending within a
  step, descending in the adjoint) -- exactly what ``step``, ``GradModel.forward_step`` and
  ``backward_step`` do;
* one simulator can hold ``n_envs`` independent environments (the reference: one per process).
"""
import numpy as np

from ..engine import Engine, EngineError
from ..scene import SceneSpec, scene_from_cfg
from .fields import FrameField, ParticleScalarField, ScalarField, ZeroOnFillGrad


class _KernelWithGrad:
    """Callable with a ``.grad`` attribute, like a Taichi kernel (``compute_grid_m_kernel.grad(f)``)."""

    def __init__(self, fwd, bwd):
        self._fwd, self.grad = fwd, bwd

    def __call__(self, *a, **k):
        return self._fwd(*a, **k)


class _GridMassField:
    """sim.grid_m: ``.fill(0)``, ``.to_torch(device)``, ``.to_numpy()``, ``.grad.from_torch(t)`` (function.py:122-152)."""

    def __init__(self, sim):
        self._sim = sim
        self._value = None
        self.grad = self

    def fill(self, v):
        self._value = None

    def _set(self, value):
        self._value = value

    def to_numpy(self, env=0):
        n = self._sim.n_grid
        return np.zeros((n, n, n), np.float32) if self._value is None else self._value[env].copy()

    def to_torch(self, device='cuda', env=0):
        import torch
        return torch.from_numpy(self.to_numpy(env)).to(device)

    def from_torch(self, t, env=None):      # grid_m.grad.from_torch(...); env: the env of a batched engine the gradient is for
        self._sim._grid_m_grad = t.detach().to(dtype=t.dtype).contiguous()
        self._sim._grid_m_grad_env = env


class MPMSimulator:
    def __init__(self, cfg, primitives=(), n_envs=1, env_cfg=None, max_env_steps=None, step_slots=None, sort=True,
                 device=0):
        """``cfg`` is the SIMULATOR node (as in the reference) -- or a prepared :class:`SceneSpec`."""
        self.primitives = primitives
        self.n_primitive = len(primitives)
        if isinstance(cfg, SceneSpec):
            scene = cfg
        else:
            from ..config import CfgNode
            full = CfgNode(dict(SIMULATOR=cfg, PRIMITIVES=[p.cfg for p in primitives], SHAPES=[], ENV=env_cfg or {}))
            scene = scene_from_cfg(full)
        self.scene = scene
        self.cfg = cfg
        self.dim, self.dtype = 3, 'float32'
        self.n_grid, self.dx, self.inv_dx, self.dt = scene.n_grid, scene.dx, scene.inv_dx, scene.dt
        self.p_vol, self.p_rho, self.p_mass = scene.p_vol, 1, scene.p_mass
        self._mu, self._lam, self._yield_stress = scene.mu, scene.lam, scene.yield_stress
        self.ground_friction, self.default_gravity, self.lower_bound = scene.ground_friction, scene.gravity, scene.lower_bound
        self.max_steps, self.substeps = scene.max_steps, scene.substeps
        self.res = (self.n_grid,) * 3
        self.num_particle_collision = 600
        self.collision_pairs = [list(p) for p in scene.pairs]
        self.num_collision = len(scene.pairs)
        self.n_envs = n_envs
        self.horizon = max_env_steps or max(1, self.max_steps // self.substeps)   # env steps the tape can hold
        self.engine = Engine(scene, n_envs=n_envs, capacity=scene.particle_capacity, max_steps=self.horizon,
                             step_slots=step_slots or 1, sort=sort, device=device)
        self.cur = 0
        self._actions = np.zeros((self.horizon, n_envs, max(1, scene.action_dim)), np.float32)
        self._grid_m_grad = None
        self._grid_m_grad_env = None
        zero = self.engine.zero_grad
        self.x, self.v = FrameField(self, 'x', zero), FrameField(self, 'v', zero)
        self.F, self.C = FrameField(self, 'F', zero), FrameField(self, 'C', zero)
        self.n_particles = ScalarField(lambda: self.engine.n_particles(0))
        self.mu = ParticleScalarField(self, 'mu', self._mu)
        self.lam = ParticleScalarField(self, 'lam', self._lam)
        self.yield_stress = ParticleScalarField(self, 'yield_stress', self._yield_stress)
        self.grid_m = _GridMassField(self)
        self.compute_grid_m_kernel = _KernelWithGrad(self._compute_grid_m, self._compute_grid_m_grad)
        if hasattr(primitives, 'primitives'):
            primitives._sim = self
            for i, p in enumerate(primitives.primitives):
                p._bind(self, i)

    # ---- frames <-> steps ----------------------------------------------------------------------------------
    def _frame_to_step(self, f):
        if f % self.substeps != 0:
            raise IndexError(f"frame {f} is not an env-step boundary (substeps={self.substeps}); the engine keeps "
                             "checkpoints at step boundaries only")
        s = f // self.substeps
        if s > self.horizon:
            raise IndexError(f"frame {f} beyond the engine horizon of {self.horizon} env steps "
                             f"(the reference silently overruns its {self.max_steps}-frame tape here)")
        return s

    def set_num_collision(self):
        return self.num_collision

    def initialize(self, n_particles):
        """mpm_simulator.py:82-97: material fill, gravity, rand_num (the engine did the latter at creation)."""
        self._n_init = n_particles
        self.engine.set_gravity(self.default_gravity)

    # ---- stepping ------------------------------------------------------------------------------------------
    def _set_action(self, s, action, env=None):
        action = np.asarray(action, dtype=np.float32).reshape(-1)
        if env is None:
            self._actions[s, :, :len(action)] = action
        else:
            self._actions[s, env, :len(action)] = action
        self.engine.set_action(s, self._actions[s, :, :self.scene.action_dim])

    def _set_tool_action(self, index, s, action_slice):
        lo = self.scene.action_dims[index]
        self._actions[s, :, lo:lo + len(action_slice)] = np.clip(action_slice, -1, 1)
        self.engine.set_action(s, self._actions[s, :, :self.scene.action_dim])

    def substep(self, s):
        self.engine.substep(s)

    def substep_grad(self, s):
        self.engine.substep_grad(s)

    def step(self, is_copy, action=None):
        """mpm_simulator.py:440-451."""
        start = 0 if is_copy else self.cur
        s = start // self.substeps
        self.cur = start + self.substeps
        if action is not None:
            self.primitives.set_action(s, self.substeps, action)
        if is_copy:
            self.engine.forward_step(0, 0, 0)       # S substeps, result copied back to frame 0
            self.cur = 0
        else:
            self.engine.forward_step(s, s + 1, s)

    def copyframe(self, source, target):
        self.engine.copy_step(self._frame_to_step(source), self._frame_to_step(target))

    # ---- io ------------------------------------------------------------------------------------------------
    def _particles(self, f, fields, env=0):
        if f % self.substeps == 0:
            return self.engine.get_particles(self._frame_to_step(f), env, fields)
        try:
            x, v, F, C = self.engine.debug_frame(f, env)
        except EngineError as e:
            raise IndexError(f"frame {f} is neither an env-step boundary (substeps={self.substeps}) nor resident in "
                             f"the engine's step-slot ring: {e}") from None
        d = dict(x=x, v=v, F=F, C=C)
        return tuple(d[k] for k in fields)

    def get_state(self, f, env=0):
        x, v, F, C = (a.astype(np.float64) for a in self._particles(f, 'xvFC', env))
        out = [x, v, F, C]
        for p in self.primitives:
            out.append(p.get_state(f, env))
        return out

    def set_state(self, f, state, env=None):
        s = self._frame_to_step(f)
        for b in (range(self.n_envs) if env is None else [env]):
            self.engine.set_particles(s, b, *[np.asarray(a, dtype=np.float32) for a in state[:4]])
        for st, p in zip(state[4:], self.primitives):
            p.set_state(f, st, env)

    def set_primitive_state(self, f, state):
        for st, p in zip(state[4:], self.primitives):
            p.set_state(f, st)

    def reset(self, x, env=None):
        for b in (range(self.n_envs) if env is None else [env]):
            self.engine.set_particles(0, b, np.asarray(x, dtype=np.float32))     # v=0, F=I, C=0
        self.cur = 0

    def get_x(self, f, env=0):
        return self._particles(f, 'x', env)[0].astype(np.float64)

    def get_v(self, f, env=0):
        return self._particles(f, 'v', env)[0].astype(np.float64)

    def get_torch_x(self, f, device='cuda', env=0):
        import torch
        n = self.engine.n_particles(env)
        x = torch.zeros((n, 3), device=device, dtype=torch.float32)
        if x.is_cuda:
            self.engine._ck(self.engine.L.dsk_get_particles(self.engine.h, self._frame_to_step(f), env,
                                                           __import__('ctypes').c_void_p(x.data_ptr()), None, None, None, 1))
            return x
        return torch.from_numpy(self._particles(f, 'x', env)[0]).to(device)

    # ---- loss helpers ----------------------------------------------------------------------------------------
    def _compute_grid_m(self, f):
        self.grid_m._set(self.engine.compute_grid_m(self._frame_to_step(f)))

    def _compute_grid_m_grad(self, f):
        g = self._grid_m_grad
        if g is None:
            raise RuntimeError("grid_m.grad.from_torch(...) must be called before compute_grid_m_kernel.grad")
        n = self.n_grid
        if hasattr(g, 'is_cuda'):
            import torch
            g = g.reshape(-1, n, n, n).float().contiguous()
            if g.shape[0] != self.n_envs:
                # a single-env gradient on a batched engine goes to ITS env only (the other envs get zeros)
                full = torch.zeros((self.n_envs, n, n, n), dtype=g.dtype, device=g.device)
                full[getattr(self, '_grid_m_grad_env', None) or 0] = g[0]
                g = full
            if not g.is_cuda:
                g = g.numpy()
        self.engine.compute_grid_m_grad(self._frame_to_step(f), g)

    def clear_and_compute_grid_m(self, f):
        self.grid_m.fill(0)
        self.compute_grid_m_kernel(f)

    def get_m(self):
        return self.grid_m.to_numpy()
