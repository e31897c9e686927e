"""Environment registry for the three DiffSkill envs (plb/envs/__init__.py:30-66), gym-free.

``make(name)`` returns a small env object with the reference's ``reset / step / taichi_env`` surface; the
dataset-backed ``MultitaskPlasticineEnv`` (cached init/target states from Google Drive) is a caller of the hot
path and is not reproduced -- initial states are the synthetic doughs of ``diffskill_b200.shapes``.
"""
import numpy as np

from ..config import load
from .scenes import SCENES


class PlasticineEnv:
    def __init__(self, name, seed=100, n_envs=1, **kw):
        from ..sim import TaichiEnv
        self.name = name
        self.cfg = load(data=SCENES[name])
        self.taichi_env = TaichiEnv(self.cfg, loss=False, n_envs=n_envs, **kw)
        self.taichi_env.initialize()
        self.action_dim = self.taichi_env.primitives.action_dim
        self.rng = np.random.RandomState(seed)
        self._init_state = self.taichi_env.get_state()
        self.horizon, self._t = 50, 0

    def seed(self, seed):
        self.rng = np.random.RandomState(seed)

    def sample_action(self):
        return self.rng.uniform(-1, 1, self.action_dim)

    def reset(self):
        self.taichi_env.set_state(**self._init_state)
        self._t = 0
        return self.taichi_env.simulator.get_x(0)

    def step(self, action):
        action = np.asarray(action).clip(-1, 1)
        self.taichi_env.step(action)
        self._t += 1
        x = self.taichi_env.simulator.get_x(0)
        if np.isnan(x).any():
            raise Exception("NaN..")                    # multitask_env.py:140-146
        return x, 0., self._t >= self.horizon, {}


def make(env_name, **kwargs):
    if env_name not in SCENES:
        raise KeyError(f"unknown env {env_name!r}; registered: {sorted(SCENES)}")
    return PlasticineEnv(env_name, **kwargs)
