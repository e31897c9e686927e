"""On-disk formats either side of the hot path.

The reference stores cached start/goal pairs as ``<cached_state_path>/init/state_<i>.xz`` -- an lzma-compressed
pickle of ``TaichiEnv.get_state()`` (``{'state': [x, v, F, C, tool_0, ...], 'softness', 'is_copy'}``,
core/diffskill/gen_init_target/state_generator.py:34-57, plb/envs/multitask_env.py:24-32,85-89) -- and
``target/target_<i>.npy`` (goal particle positions).  The Google-Drive dataset is unavailable offline; these helpers
read/write the same layout so a real dataset drops in, and ``generate_synthetic`` fills a directory with synthetic
pairs built like the reference's generators (settle under zero actions, then a goal shape).
"""
import lzma
import os
import pickle

import numpy as np


def save_state(path, state):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with lzma.open(path, 'wb') as f:
        pickle.dump(state, f, protocol=4)


def load_state(path):
    with lzma.open(path, 'rb') as f:
        return pickle.load(f)


def init_path(root, i):
    return os.path.join(root, 'init', f'state_{i}.xz')


def target_path(root, i):
    return os.path.join(root, 'target', f'target_{i}.npy')


def save_pair(root, i, state, target_x):
    save_state(init_path(root, i), state)
    os.makedirs(os.path.join(root, 'target'), exist_ok=True)
    np.save(target_path(root, i), np.asarray(target_x, dtype=np.float64))


def load_pair(root, i):
    return load_state(init_path(root, i)), np.load(target_path(root, i))


def list_pairs(root):
    d = os.path.join(root, 'init')
    if not os.path.isdir(d):
        return []
    ids = [int(f[len('state_'):-3]) for f in os.listdir(d) if f.startswith('state_') and f.endswith('.xz')]
    return sorted(i for i in ids if os.path.exists(target_path(root, i)))


def generate_synthetic(env, root, n_pairs, settle_steps=10, seed=0):
    """Synthetic start/goal pairs through the engine: reset, `settle_steps` zero actions (the dough drops onto the
    floor / tools as in gathermove_generator_V2.py:30-32), save the state; goal = the settled dough flattened and
    shifted (a deterministic stand-in for the reference's scripted goals)."""
    rng = np.random.RandomState(seed)
    te = env.taichi_env
    for i in range(n_pairs):
        env.reset()
        for _ in range(settle_steps):
            env.step(np.zeros(env.action_dim))
        st = te.get_state()
        x = st['state'][0]
        c = x.mean(0)
        shift = np.array([rng.uniform(-0.1, 0.1), 0.0, rng.uniform(-0.05, 0.05)])
        goal = (x - c) * np.array([1.3, 0.6, 1.3]) + c + shift
        goal[:, 1] -= goal[:, 1].min() - x[:, 1].min()
        save_pair(root, i, st, goal)
    return list_pairs(root)
