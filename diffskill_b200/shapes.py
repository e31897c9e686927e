"""Synthetic dough generators (numpy restatement of plb/engine/shapes/shape_maker.py).

The Google-Drive start/goal dataset is unavailable offline, so initial particle
clouds are generated exactly as ``Shapes`` does: box :81-93, sphere :104-116,
scatter :37-51, particle count rule :56.  The reference draws from numpy's
global RNG; every generator here takes an explicit ``rng`` / seed instead.
"""
import numpy as np


def get_n_particles(volume):
    return max(int(volume / (0.1 ** 3) * 30000), 1)     # shape_maker.py:56


def make_box(init_pos, width, n_particles=None, rng=None):
    rng = rng or np.random.RandomState(0)
    width = np.array([width] * 3 if isinstance(width, float) else width, dtype=np.float64)
    if n_particles is None:
        n_particles = get_n_particles(np.prod(width))
    return (rng.random_sample((n_particles, 3)) * 2 - 1) * (0.5 * width) + np.array(init_pos)


def make_sphere(init_pos, radius, n_particles=None, rng=None):
    rng = rng or np.random.RandomState(0)
    if n_particles is None:
        n_particles = get_n_particles((radius ** 3) * 4 * np.pi / 3)
    p = rng.normal(size=(n_particles, 3))
    p /= np.linalg.norm(p, axis=-1, keepdims=True)
    u = rng.random_sample(size=(n_particles, 1)) ** (1. / 3)
    return p * u * radius + np.array(init_pos)[:3]


def make_scatter(pos_min, pos_max, seed):
    rng = np.random.RandomState(seed)
    N, multiply = 40, 50
    cols = [rng.uniform(lo, hi, size=N).reshape(N, 1) for lo, hi in zip(pos_min, pos_max)]
    centres = np.hstack(cols)
    noise = rng.normal(0, scale=0.004, size=N * 3 * multiply).reshape(multiply, N, 3)
    noise[:, :, 1] = noise[:, :, 1] * 0.2 + 0.1
    return (centres.reshape(1, N, 3) + noise).reshape(multiply * N, 3)


class Shapes:
    """``Shapes(cfg).get()`` -> (particles[N,3], colors[N]) as shape_maker.py:13-169 (box/sphere/scatter)."""

    def __init__(self, cfg, seed=0):
        self.objects, self.colors = [], []
        rng = np.random.RandomState(seed)
        for i in cfg:
            kw = {k: v for k, v in dict(i).items() if k != 'shape'}
            color = kw.pop('color', None)
            if i['shape'] == 'box':
                p = make_box(kw['init_pos'], kw['width'], kw.get('n_particles'), rng)
            elif i['shape'] == 'sphere':
                p = make_sphere(kw['init_pos'], kw['radius'], kw.get('n_particles'), rng)
            elif i['shape'] == 'scatter':
                p = make_scatter(kw['pos_min'], kw['pos_max'], kw['seed'])
            else:
                raise NotImplementedError(f"Shape {i['shape']} is not supported!")
            self.objects.append(p)
            c = np.zeros(len(p), np.int32)
            c[:] = color if isinstance(color, int) else 127
            self.colors.append(c)

    def get(self):
        assert len(self.objects) > 0, "please add at least one shape into the scene"
        return np.concatenate(self.objects), np.concatenate(self.colors)
