"""Synthetic dough generators (numpy restatement of plb/engine/shapes/shape_maker.py).

The Google-Drive start/goal dataset is unavailable offline, so initial particle
clouds are generated exactly as ``Shapes`` does: box :81-93, multibox :72-80, sphere :104-116,
multisphere :93-103, capsule :118-147, cylinder :149-166, scatter :37-51, the optional rotation
of an object about its centroid :58-63, particle count rule :56.  The reference draws from numpy's
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


def quat_to_mat(q):
    """Rotation matrix of a (w, x, y, z) quaternion (what transforms3d.quaternions.quat2mat returns, shape_maker.py:60)."""
    w, x, y, z = [float(v) for v in q]
    n = w * w + x * x + y * y + z * z
    s = 0.0 if n < np.finfo(float).eps else 2.0 / n
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ, xX, xY, xZ, yY, yZ, zZ = w * X, w * Y, w * Z, x * X, x * Y, x * Z, y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY],
                     [xY + wZ, 1.0 - (xX + zZ), yZ - wX],
                     [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def rotate_about_centroid(p, init_rot):
    """add_object's init_rot (shape_maker.py:58-63): rotate the cloud about its mean."""
    if init_rot is None:
        return p
    origin = p.mean(axis=0)
    return (p - origin) @ quat_to_mat(init_rot).T + origin


def make_cylinder(init_pos, radius, height, n_particles=None, rng=None):
    """Axis along z (shape_maker.py:149-166).  The particle count follows the reference's rule, which uses the SPHERE volume
    of the radius for capsules and cylinders alike (:151-155)."""
    rng = rng or np.random.RandomState(0)
    if n_particles is None:
        n_particles = get_n_particles((radius ** 3) * 4 * np.pi / 3)
    r = radius * np.sqrt(rng.random_sample((n_particles, 1)))
    theta = rng.random_sample((n_particles, 1)) * 2 * np.pi
    h = rng.random_sample((n_particles, 1)) * height - height / 2.
    return np.hstack([np.cos(theta) * r, np.sin(theta) * r, h]) + np.array(init_pos)


def make_capsule(init_pos, radius, height, n_particles=None, rng=None):
    """Two half balls pushed apart along z plus a cylinder between them (shape_maker.py:118-147), split by volume."""
    rng = rng or np.random.RandomState(0)
    if n_particles is None:
        n_particles = get_n_particles((radius ** 3) * 4 * np.pi / 3)
    v_sphere = (radius ** 3) * 4 * np.pi / 3
    v_cylinder = (radius ** 2) * np.pi * height
    n1 = int(n_particles * v_sphere / (v_sphere + v_cylinder))
    n2 = n_particles - n1
    p = rng.normal(size=(n1, 3))
    p /= np.linalg.norm(p, axis=-1, keepdims=True)
    p = p * rng.random_sample(size=(n1, 1)) ** (1. / 3) * radius
    p[:, 2] += np.sign(p[:, 2]) * height / 2.
    p += np.array(init_pos)[:3]
    return np.vstack([p, make_cylinder(init_pos, radius, height, n2, rng)])


def make_multibox(all_pos, all_width, all_rot=None, rng=None):
    """Several boxes sharing one particle budget in proportion to their volumes (shape_maker.py:72-80; capped at 30 000)."""
    rng = rng or np.random.RandomState(0)
    volumes = [float(np.prod(w)) for w in all_width]
    total = sum(volumes)
    n = min(30000, get_n_particles(total))
    all_rot = all_rot or [None] * len(all_pos)
    return [rotate_about_centroid(make_box(pos, w, int(n * v / total), rng), rot)
            for pos, w, rot, v in zip(all_pos, all_width, all_rot, volumes)]


def make_multisphere(all_pos, all_r, rng=None):
    """Several spheres sharing one particle budget in proportion to their volumes (shape_maker.py:93-103)."""
    rng = rng or np.random.RandomState(0)
    volumes = [(r ** 3) * 4 * np.pi / 3 for r in all_r]
    total = sum(volumes)
    n = get_n_particles(total)
    return [make_sphere(pos, r, int(n * v / total), rng) for pos, r, v in zip(all_pos, all_r, volumes)]


def make_scatter(pos_min, pos_max, seed):
    rng = np.random.RandomState(seed)
    N, multiply = 40, 50
    cols = [rng.uniform(lo, hi, size=N).reshape(N, 1) for lo, hi in zip(pos_min, pos_max)]
    centres = np.hstack(cols)
    noise = rng.normal(0, scale=0.004, size=N * 3 * multiply).reshape(multiply, N, 3)
    noise[:, :, 1] = noise[:, :, 1] * 0.2 + 0.1
    return (centres.reshape(1, N, 3) + noise).reshape(multiply * N, 3)


COLORS = [(127 << 16) + 127, (127 << 8), 127, 127 << 16]   # shape_maker.py:5-10


class Shapes:
    """``Shapes(cfg).get()`` -> (particles[N,3], colors[N]) as shape_maker.py:13-169: box, multibox, sphere, multisphere,
    capsule, cylinder, scatter; string-valued entries are evaluated (:24) and `init_rot` rotates an object about its
    centroid (:58-63)."""

    def __init__(self, cfg, seed=0):
        self.objects, self.colors = [], []
        rng = np.random.RandomState(seed)
        for i in cfg:
            kw = {k: (eval(v) if isinstance(v, str) else v) for k, v in dict(i).items() if k != 'shape'}
            color = kw.pop('color', None)
            rot = kw.pop('init_rot', None)
            shape = i['shape']
            if shape == 'box':
                parts = [rotate_about_centroid(make_box(kw['init_pos'], kw['width'], kw.get('n_particles'), rng), rot)]
            elif shape == 'multibox':
                parts = make_multibox(kw['all_pos'], kw['all_width'], kw.get('all_rot'), rng)
            elif shape == 'sphere':
                parts = [rotate_about_centroid(make_sphere(kw['init_pos'], kw['radius'], kw.get('n_particles'), rng), rot)]
            elif shape == 'multisphere':
                parts = make_multisphere(kw['all_pos'], kw['all_r'], rng)
            elif shape == 'capsule':
                parts = [rotate_about_centroid(make_capsule(kw['init_pos'], kw['radius'], kw['height'], kw.get('n_particles'), rng), rot)]
            elif shape == 'cylinder':
                parts = [rotate_about_centroid(make_cylinder(kw['init_pos'], kw['radius'], kw['height'], kw.get('n_particles'), rng), rot)]
            elif shape == 'scatter':
                parts = [make_scatter(kw['pos_min'], kw['pos_max'], kw['seed'])]
            else:
                raise NotImplementedError(f"Shape {i['shape']} is not supported!")
            for p in parts:
                c = np.zeros(len(p), np.int32)
                # add_object: an int colour as given, otherwise the palette entry of the object's index (:64-68)
                c[:] = color if isinstance(color, int) else COLORS[len(self.objects) % len(COLORS)]
                self.objects.append(p)
                self.colors.append(c)

    def get(self):
        assert len(self.objects) > 0, "please add at least one shape into the scene"
        return np.concatenate(self.objects), np.concatenate(self.colors)
