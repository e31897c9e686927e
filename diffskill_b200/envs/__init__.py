"""Environment registry for the three DiffSkill envs (plb/envs/__init__.py:30-66), gym-free.

``make(name)`` returns a small env object with the reference's ``reset / step / taichi_env`` surface over the synthetic
doughs of ``diffskill_b200.shapes``; ``MultitaskPlasticineEnv`` below is the dataset-backed env (cached init/target states in
the reference's on-disk layout, written here by ``envs.dataset.generate_synthetic`` since the Google-Drive set is offline).
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


class MultitaskPlasticineEnv:
    """plb/envs/multitask_env.py:34-200 without gym: the dataset-backed env.  `cached_state_path` holds `init/state_<i>.xz`
    (lzma pickle of TaichiEnv.get_state()) and `target/target_<i>.npy` (goal particles) -- diffskill_b200.envs.dataset reads and
    writes that layout; `target/target_imgs.npy` (rendered goals) is loaded when present.  reset(init_v, target_v,
    target_cfg_modifier, contact_loss_mask) / step / get_state / set_state / get_primitive_state as in the reference; the
    scripted `primitive_reset_to` policies are callers' code and are not mirrored."""

    def __init__(self, name, cached_state_path=None, version=None, nn=False, loss=True, return_dist=False,
                 generating_cached_state=False, device=None, **kw):
        import glob
        import os
        from ..sim import TaichiEnv
        self.name = self.cfg_path = name
        full = self._load_cfg()
        if cached_state_path is not None:
            full.ENV.cached_state_path = cached_state_path
        self._root = full.ENV.cached_state_path
        self.taichi_env = TaichiEnv(full, nn, loss=loss, return_dist=return_dist, **kw)
        if device is not None:
            self.taichi_env.device = device
        self.taichi_env.initialize(full, target_path=None)
        self.generating_cached_state = generating_cached_state
        self.cfg = full.ENV
        if not generating_cached_state:
            self.num_inits = len(glob.glob(os.path.join(self._root, 'init', 'state_*.xz')))
            self.num_targets = len(glob.glob(os.path.join(self._root, 'target', 'target_[0-9]*.npy')))
            imgs = os.path.join(self._root, 'target', 'target_imgs.npy')
            self.target_imgs = np.load(imgs) if os.path.exists(imgs) else None
            ids = sorted(int(os.path.basename(p)[len('target_'):-4]) for p in
                         glob.glob(os.path.join(self._root, 'target', 'target_[0-9]*.npy')))
            self.target_pcs = [np.load(os.path.join(self._root, 'target', f'target_{i}.npy')) for i in ids]   # natsorted order
        self.taichi_env.set_copy(True)
        self._init_state = self.taichi_env.get_state()
        self.action_dim = self.taichi_env.primitives.action_dim
        self.reset()

    def _load_cfg(self):
        full = load(data=SCENES[self.name])
        if getattr(self, '_root', None) is not None:
            full.ENV.cached_state_path = self._root
        return full

    @property
    def action_dims(self):
        return self.taichi_env.primitives.action_dims

    def reset(self, init_v=None, target_v=None, target_cfg_modifier=None, contact_loss_mask=None):
        import os
        import torch
        from .dataset import load_state
        full = self._load_cfg()                               # the reference reloads the cfg on every reset
        self.cfg = full.ENV
        if target_cfg_modifier is not None:
            target_cfg_modifier(full)
        te = self.taichi_env
        target_path = None
        if not self.generating_cached_state:
            if init_v is None:
                assert target_v is None
                init_v, target_v = np.random.randint(0, self.num_inits), np.random.randint(0, self.num_targets)
            self.init_v, self.target_v = init_v, target_v
            self.target_img = None if self.target_imgs is None else self.target_imgs[target_v]
            self.target_pc = self.target_pcs[target_v]
            target_path = os.path.join(self._root, 'target', f'target_{target_v}.npy')
        te.initialize(full, target_path=target_path)
        te.set_copy(True)
        if not self.generating_cached_state:
            self._init_state = load_state(os.path.join(self._root, 'init', f'state_{init_v}.xz'))
            te.set_state(**self._init_state)
        self._n_observed_particles = self.cfg.n_observed_particles
        self._recorded_actions = []
        te.set_init_emd()
        te.contact_loss_mask = torch.zeros(len(te.primitives), device=te.device)
        if isinstance(contact_loss_mask, (int, float)):
            te.contact_loss_mask[int(contact_loss_mask)] = 1.
        elif isinstance(contact_loss_mask, list):
            for i in range(len(te.primitives)):
                te.contact_loss_mask[i] = contact_loss_mask[i]
        return self._get_obs()

    def reset_primitive(self):
        self.taichi_env.set_primitive_state(**self._init_state)

    @staticmethod
    def state_to_vec(d):
        return np.concatenate([np.asarray(d[k]).flatten() for k in ('particles', 'tool_state', 'tool_particles')])

    def _get_obs(self, t=0):
        particles, tool_state = self.taichi_env.get_obs(t, device='cpu')
        idx = np.arange(0, min(1000, particles.shape[0]))     # the reference hard-codes the first 1000 particles
        return self.state_to_vec({'particles': particles[idx].numpy(), 'tool_state': tool_state.numpy(),
                                  'tool_particles': self.taichi_env.get_tool_particles(0)})

    def step(self, action):
        action = np.clip(action, -1., 1.)
        self.taichi_env.step(action)
        r, info = self.taichi_env.get_reward_and_info()
        self._recorded_actions.append(action)
        obs = self._get_obs()
        if np.isnan(obs).any() or np.isnan(r):
            import datetime
            import pickle
            with open(f'{self.cfg_path}_nan_action_{datetime.datetime.now()}', 'wb') as f:
                pickle.dump(self._recorded_actions, f)
            raise Exception("NaN..")
        return obs, r, False, info

    def render(self, mode='human', *args, **kwargs):
        return self.taichi_env.render(mode, *args, **kwargs)

    def get_state(self):
        return self.taichi_env.get_state()

    def set_state(self, state):
        self.taichi_env.set_state(**state)

    def get_primitive_state(self):
        return [i.get_state(0) for i in self.taichi_env.primitives]
