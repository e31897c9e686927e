"""Legacy PlasticineLab grid loss (plb/engine/losses/loss.py:8-369), SURVEY.md section 8(f) row 4.

Same class, same methods, same bookkeeping (`reset`, `compute_loss` -> reward / incremental_iou, `get_state/set_state` for CEM).
The two particle-sized passes are the engine's CUDA kernels behind the C ABI -- the mass-only P2G `dsk_compute_grid_m` and its
adjoint (mpm_simulator.py:456-471), the per-particle tool SDF `dsk_compute_min_dist` and its adjoint (function.py:79-88).  The
reductions in between (|m - target|, sdf * m, the hard / soft minimum over particles, IoU) are a handful of torch ops on the
tensors those kernels fill, on the same device; their derivatives -- what Taichi's autodiff generates for loss.py:134-182 -- come
from torch.autograd on exactly these expressions and are handed to the adjoint kernels, which accumulate into the engine's
x.grad / tool-pose adjoints of that step (the reference's `compute_loss_kernel_grad`, loss.py:262-289).

Differences from the reference, by construction of the engine: frames are env-step boundaries (`MPMSimulator._frame_to_step`);
the engine may hold several envs, the loss belongs to one of them (`env`, default 0).  Not verifiable offline: the adjoint
Taichi 0.7.26 gives `ti.atomic_min` (hard contact, loss.py:146-151); here the minimum's adjoint goes to the arg-min particle
(the subgradient torch.min uses).  The reference's default is the hard minimum (`soft_contact=False`, default_config.py:71).
"""
import os

import numpy as np
import torch

INF = 1000.0          # loss.py:36


def target_sdf_sweep(density, sdf_copy, nearest_copy, dx, inf=INF):
    """One launch of `update_target_sdf` (loss.py:105-125): Jacobi sweep over the 6x6x6 neighbourhood (offsets -3..2, x
    outermost, the first strictly smaller distance wins), reading the *_copy fields of the previous sweep.
    density, sdf_copy: [n,n,n]; nearest_copy: [n,n,n,3]; returns (sdf, nearest) of this sweep."""
    n = density.shape[0]
    dev, dt = density.device, torch.float32
    ax = torch.arange(n, device=dev, dtype=dt) * dx
    pos = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), -1)                    # grid_pos = I * dx
    sdf = torch.full((n, n, n), inf, device=dev, dtype=dt)
    nearest = nearest_copy.clone()          # nodes no sweep reaches keep what the field held (the reference never clears it)
    P = 3
    sp = torch.nn.functional.pad(sdf_copy, (P, P) * 3, value=inf)
    npad = torch.nn.functional.pad(nearest_copy, (0, 0) + (P, P) * 3, value=0.)
    for ox in range(-3, 3):
        for oy in range(-3, 3):
            for oz in range(-3, 3):
                if ox == 0 and oy == 0 and oz == 0:
                    continue
                s = sp[P + ox:P + ox + n, P + oy:P + oy + n, P + oz:P + oz + n]
                q = npad[P + ox:P + ox + n, P + oy:P + oy + n, P + oz:P + oz + n]
                d = pos - q
                dist = torch.sqrt((d * d).sum(-1) + 1e-8)                                 # self.norm, loss.py:101-103
                take = (s < inf) & (dist < sdf)
                sdf = torch.where(take, dist, sdf)
                nearest = torch.where(take[..., None], q, nearest)
    solid = density > 1e-4
    sdf = torch.where(solid, torch.zeros_like(sdf), sdf)
    nearest = torch.where(solid[..., None], pos, nearest)
    return sdf, nearest


def soft_weight(d):
    return 1 / (1 + d * d * 10000)          # loss.py:135-137


class _Scalar:
    """`loss.sdf_weight[None] = v` style 0-d field."""

    def __init__(self, v=0.):
        self.v = float(v)

    def __getitem__(self, _):
        return self.v

    def __setitem__(self, _, v):
        self.v = float(v)


class Loss:
    def __init__(self, cfg, sim, env=0, device=None):
        self.cfg, self.sim, self.env = cfg, sim, env
        self.engine = sim.engine
        self.dtype, self.res, self.n_grid, self.dx, self.dim = 'float32', (sim.n_grid,) * 3, sim.n_grid, sim.dx, 3
        self.device = torch.device(device if device is not None else ('cuda' if torch.cuda.is_available() else 'cpu'))
        # only the movable tools take part in the contact term (loss.py:22-26); their distance columns in the engine's
        # [B, capacity, ncols] table: two per 8-dof tool (function.py:23-27), one otherwise
        self.primitives, self._cols, k = [], [], 0
        for p in sim.primitives:
            w = getattr(p, 'dist_cols', 2 if p.state_dim == 8 else 1)
            if p.action_dim > 0:
                self.primitives.append(p)
                self._cols.append((k, k + w))
            k += w
        n = self.n_grid
        z = dict(device=self.device, dtype=torch.float32)
        self.target_density = torch.zeros((n, n, n), **z)
        self.target_sdf = torch.zeros((n, n, n), **z)
        self.nearest_point = torch.zeros((n, n, n, 3), **z)
        self.target_sdf_copy = torch.zeros((n, n, n), **z)
        self.nearest_point_copy = torch.zeros((n, n, n, 3), **z)
        self.inf = INF
        self.sdf_weight, self.density_weight, self.chamfer_weight = _Scalar(), _Scalar(), _Scalar()
        self.contact_weight = [0.] * len(sim.primitives)
        self.soft_contact_loss = False
        self.enable_target_update = True
        self._target_iou = None
        self.loss = self.sdf_loss = self.density_loss = self.contact_loss = 0.
        self.min_dist = [0.] * len(self.primitives)
        self._start_loss = self._init_iou = self._last_loss = 0.
        self._iou = 0.
        self.sweeps = 0

    # ---- target ----------------------------------------------------------------------------------------------------
    def load_target_density(self, path=None, grids=None):         # loss.py:57-74
        if grids is None and (path is None or len(path) == 0):
            return              # the reference's empty default `target_path: ''` has no grid to load (default_config.py:70)
        if path is not None and len(path) > 0:
            grids = np.load(path if os.path.isabs(path) else os.path.join(os.path.dirname(os.path.abspath(__file__)), '../../', path))
        else:
            grids = np.array(grids)
        self.target_density = torch.as_tensor(grids, dtype=torch.float32, device=self.device).contiguous()
        self.update_target()
        self._target_iou = self._iou_of(self.target_density)
        idxes = (np.array(np.where(grids > 1e-4)).transpose() * self.dx).astype(grids.dtype)
        self.num_particle_target = idxes.shape[0]
        self.particle_target = idxes

    def initialize(self):                                          # loss.py:76-85
        self.set_weights_only(self.cfg.weight.sdf, self.cfg.weight.density, self.cfg.weight.contact, self.cfg.soft_contact, 0.)
        self.load_target_density(self.cfg.target_path)

    def set_weights_only(self, sdf, density, contact, is_soft_contact, chamfer):
        self.sdf_weight[None], self.density_weight[None], self.chamfer_weight[None] = sdf, density, chamfer
        self.contact_weight = [float(contact)] * len(self.contact_weight)
        self.soft_contact_loss = bool(is_soft_contact)

    def set_weights(self, sdf, density, contact, is_soft_contact, chamfer):   # loss.py:87-94
        self.set_weights_only(sdf, density, contact, is_soft_contact, chamfer)
        self.reset()

    def set_target_update(self, flag):
        self.enable_target_update = flag

    def update_target(self):                                       # loss.py:130-134
        """2*n_grid Jacobi sweeps in the reference; a sweep is a pure function of the *_copy fields, so the loop stops at the
        first sweep that changes nothing (the remaining ones would be identities)."""
        self.target_sdf_copy.fill_(self.inf)
        self.sweeps = 0
        if not self.enable_target_update:
            return
        for _ in range(self.n_grid * 2):
            sdf, nearest = target_sdf_sweep(self.target_density, self.target_sdf_copy, self.nearest_point_copy, self.dx, self.inf)
            same = torch.equal(sdf, self.target_sdf_copy) and torch.equal(nearest, self.nearest_point_copy)
            self.target_sdf, self.nearest_point = sdf, nearest
            self.target_sdf_copy, self.nearest_point_copy = sdf.clone(), nearest.clone()
            self.sweeps += 1
            if same:
                break

    # ---- engine passes -----------------------------------------------------------------------------------------------
    def _grid_mass(self, f):
        """grid_mass.fill(0); compute_grid_mass(f) (loss.py:246-247): the engine's mass-only P2G."""
        eng, n = self.engine, self.n_grid
        m = torch.zeros((eng.B, n, n, n), device=self.device, dtype=torch.float32)
        eng.compute_grid_m(self.sim._frame_to_step(f), m)
        return m

    def _tool_sdf(self, f):
        eng = self.engine
        d = torch.zeros((eng.B, eng.capacity, eng.ncols), device=self.device, dtype=torch.float32)
        eng.compute_min_dist(self.sim._frame_to_step(f), d)
        return d

    # ---- the terms (loss.py:139-182), as differentiable torch expressions on the kernels' outputs -----------------------------
    def _contact_terms(self, cols):
        """cols: [n, ncols] per-particle SDF columns of this env -> list of min_dist per movable tool."""
        out = []
        for l, r in self._cols:
            s = cols[:, l] if r - l == 1 else torch.minimum(cols[:, l], cols[:, l + 1])   # Gripper.sdf = min of the two jaws
            d = torch.clamp(s, min=0.)                                                     # d_ij = max(sdf, 0)
            if self.soft_contact_loss:
                w = soft_weight(d)
                out.append((d * w).sum() / w.sum())                                        # loss.py:139-158
            else:
                out.append(torch.minimum(d.min(), torch.tensor(100000., device=d.device)) if d.numel() else
                           torch.tensor(100000., device=d.device))                         # loss.py:146-151, 243-244
        return out

    def _terms(self, m, cols):
        density = (m - self.target_density).abs().sum()                                    # loss.py:168-171
        sdf = (self.target_sdf * m).sum()                                                  # loss.py:173-176
        mins = self._contact_terms(cols) if self.primitives else []
        w = [self.contact_weight[i] for i in range(len(mins))]                             # loss.py:160-163: weight i of tool i
        contact = sum((wi * md ** 2 for wi, md in zip(w, mins)), torch.zeros((), device=m.device))
        total = contact + density * self.density_weight[None] + sdf * self.sdf_weight[None]   # loss.py:210-215
        return total, contact, density, sdf, mins

    def compute_loss_kernel(self, f):                              # loss.py:238-260
        e = self.env
        n = self.engine.n_particles(e)
        m = self._grid_mass(f)
        self.grid_mass = m[e]
        cols = self._tool_sdf(f)[e, :n] if self.primitives else torch.zeros((n, 0), device=self.device)
        with torch.no_grad():
            total, contact, density, sdf, mins = self._terms(m[e], cols)
        self.contact_loss, self.density_loss, self.sdf_loss = float(contact), float(density), float(sdf)
        self.min_dist = [float(v) for v in mins]
        self.loss += float(total)                                   # sum_up_loss_kernel accumulates until clear_loss

    def compute_loss_kernel_grad(self, f, loss_grad=1.0):          # loss.py:262-289
        """Adjoint of compute_loss_kernel(f) for d(total)/d(loss) = loss_grad: accumulates into the engine's particle and
        tool-pose adjoints of that step (what the Taichi tape does when `loss.grad[None] = 1`)."""
        e, eng = self.env, self.engine
        n = eng.n_particles(e)
        step = self.sim._frame_to_step(f)
        m = self._grid_mass(f)
        allcols = self._tool_sdf(f) if self.primitives else None
        me = m[e].clone().requires_grad_(True)
        ce = (allcols[e, :n].clone() if allcols is not None else torch.zeros((n, 0), device=self.device)).requires_grad_(True)
        total = self._terms(me, ce)[0]
        gm, gc = torch.autograd.grad(total * loss_grad, [me, ce], allow_unused=True)
        if self.primitives and gc is not None:
            g = torch.zeros_like(allcols)
            g[e, :n] = gc
            eng.compute_min_dist_grad(step, g.contiguous())
        G = torch.zeros_like(m)
        G[e] = gm
        eng.compute_grid_m_grad(step, G.contiguous())

    # ---- IoU (loss.py:291-318) ------------------------------------------------------------------------------------------------
    def _iou_of(self, m):
        t = self.target_density
        ma, mb = torch.clamp(m.max(), min=0.), torch.clamp(t.max(), min=0.)      # atomic_max into a zero-initialised local
        I = (m * t).sum() / ma / mb
        U = m.sum() / ma + t.sum() / mb
        return float(I / (U - I))

    def iou(self):
        self._iou = self._iou_of(self.grid_mass)

    @staticmethod
    def iou2(a, b):
        I = np.sum(a * b)
        return I / (np.sum(a) + np.sum(b) - I)

    # ---- bookkeeping (loss.py:320-369) ------------------------------------------------------------------------------------
    def clear_loss(self):
        self.loss = 0.

    def _extract_loss(self, f):
        self.compute_loss_kernel(f)
        self.iou()
        return {'loss': self.loss, 'contact_loss': self.contact_loss, 'density_loss': self.density_loss,
                'sdf_loss': self.sdf_loss, 'iou': self._iou, 'target_iou': self._target_iou}

    def reset(self):
        self.clear_loss()
        info = self._extract_loss(0)
        self._start_loss, self._init_iou, self._last_loss = info['loss'], info['iou'], 0

    def compute_loss(self, f):
        info = self._extract_loss(f)
        r = self._start_loss - (info['loss'] - self._last_loss)
        cur = info['loss'] - self._last_loss
        self._last_loss = info['loss']
        info['reward'] = r
        info['incremental_iou'] = max(min((info['iou'] - self._init_iou) / (info['target_iou'] - self._init_iou), 1), 0)
        info['loss'] = cur
        return info

    def clear(self):
        self.clear_loss()
        self._last_loss = 0

    def get_state(self):
        return {'_start_loss': self._start_loss, '_last_loss': self._last_loss, '_init_iou': self._init_iou}

    def set_state(self, _start_loss, _last_loss, _init_iou):
        self._start_loss, self._last_loss, self._init_iou = _start_loss, _last_loss, _init_iou
