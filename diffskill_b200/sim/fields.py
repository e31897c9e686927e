"""Field-style proxies for the Taichi field accesses the reference's callers use directly.

The reference's callers poke Taichi fields: ``sim.x.grad.fill(0)``,
``sim.n_particles[None]``, ``sim.yield_stress.fill(v)``, ``prim.friction[None] = 10.``,
``prim.xyz_limit[1] = ...``, ``prim.gap[f]``, ``sim.grid_m.to_torch(device)`` ...  Device state is
owned by the engine, so these objects forward to C-ABI calls instead of exposing memory.
"""
import numpy as np


class ZeroOnFillGrad:
    """``field.grad.fill(0)`` of GradModel.reset (function.py:44-61): all adjoints are cleared together."""

    def __init__(self, zero_fn):
        self._zero = zero_fn

    def fill(self, value):
        if value != 0:
            raise NotImplementedError("adjoint fields can only be filled with 0 (GradModel.reset semantics)")
        self._zero()


class ScalarField:
    """0-D field: ``f[None]`` get/set (e.g. prim.friction, prim.softness, sim.n_particles)."""

    def __init__(self, getter, setter=None):
        self._get, self._set = getter, setter

    def __getitem__(self, idx):
        return self._get()

    def __setitem__(self, idx, value):
        if self._set is None:
            raise AttributeError("read-only field")
        self._set(value)

    def fill(self, value):
        self[None] = value


class ParticleScalarField:
    """Per-particle material field (sim.mu / lam / yield_stress, mpm_simulator.py:34-36)."""

    def __init__(self, sim, name, default):
        self._sim, self._name = sim, name
        self._host = None
        self._default = default

    def _values(self):
        n = self._sim.n_particles[None]
        if self._host is None or len(self._host) != n:
            self._host = np.full(n, self._default, np.float32)
        return self._host

    def fill(self, value):
        self._default = float(value)
        self._host = None
        self._push()

    def from_numpy(self, arr):
        self._host = np.asarray(arr, dtype=np.float32)[:self._sim.n_particles[None]].copy()
        self._push()

    def to_numpy(self):
        return self._values().copy()

    def __getitem__(self, i):
        return float(self._values()[i])

    def __setitem__(self, i, v):
        self._values()[i] = v
        self._push()

    def _push(self):
        if self._sim.engine is not None and self._sim.n_particles[None] > 0:
            for b in range(self._sim.n_envs):
                self._sim.engine.set_material(b, **{self._name: self._values()})


class FrameField:
    """sim.x / v / F / C: ``.to_numpy()`` of the whole tape is not available (checkpoints only);
    ``field[f]``-style reads go through MPMSimulator.get_state.  Only what callers use is provided."""

    def __init__(self, sim, name, zero_fn):
        self._sim, self._name = sim, name
        self.grad = ZeroOnFillGrad(zero_fn)

    def to_numpy(self, f=0):
        st = self._sim.get_state(f)
        return {'x': st[0], 'v': st[1], 'F': st[2], 'C': st[3]}[self._name]


class ToolVectorField:
    """prim.xyz_limit (shape (2,) of vec3): ``prim.xyz_limit[1] = (..)`` / ``prim.xyz_limit[0]``."""

    def __init__(self, prim):
        self._p = prim

    def __getitem__(self, i):
        return np.array(self._p._get_limit(i))

    def __setitem__(self, i, v):
        self._p._set_limit(i, v)

    def from_numpy(self, arr):
        self[0], self[1] = arr[0], arr[1]


class ToolFrameField:
    """prim.position / rotation / gap etc.: ``field[f]`` reads at a step boundary, ``.grad.fill(0)``."""

    def __init__(self, prim, sl, zero_fn):
        self._p, self._sl = prim, sl
        self.grad = ZeroOnFillGrad(zero_fn)

    def __getitem__(self, f):
        v = self._p.get_state(f)[self._sl]
        return float(v[0]) if len(v) == 1 else v
