"""TEST INFRASTRUCTURE: an independent restatement of the reference's substep in plain torch (float64), differentiated by
torch.autograd -- the second implementation SURVEY.md section 8c item 2 asks for to pin the C++ oracle.

It is written from the reference files directly (cited per function) and from the RAW scene configuration (the restated YAML
dictionaries of diffskill_b200.envs.scenes plus the defaults of plb/config/default_config.py as merged by
diffskill_b200.config.load).  It deliberately imports NONE of the derived constants of diffskill_b200/scene.py, which the
oracle and the engine share: n_grid, dx, dt, substeps, p_vol, mu, lam, the collision pairs ... are re-derived here from
mpm_simulator.py:19-39.  Gradients come from torch.autograd (with the reference's hand-written backward_svd as a custom
Function, mpm_simulator.py:131-156), not from the oracle's tape AD and not from the engine's hand-derived adjoints.

Scope: the tools of the three DiffSkill envs (RollingPinExt, Box, Gripper, Knife), dense grid, no tool-tool collision
projection (the tests check on the oracle that none fires in the compared rollouts).
"""
import math

import torch

DT = torch.float64


def vec(*a):
    return torch.tensor(a, dtype=DT)


# ---- plb/engine/primitive/utils.py ----------------------------------------------------------------------------------------
def qrot(rot, v):                                        # utils.py:9-15; v [..., 3]
    qvec = rot[1:]
    uv = torch.cross(qvec.expand_as(v), v, dim=-1)
    uuv = torch.cross(qvec.expand_as(v), uv, dim=-1)
    return v + 2 * (rot[0] * uv + uuv)


def qmul(q, r):                                          # utils.py:23-31 (normalises)
    w = r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3]
    x = r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2]
    y = r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1]
    z = r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0]
    out = torch.stack([w, x, y, z])
    return out / torch.sqrt(out.dot(out))


def w2quat(aa):                                          # utils.py:34-47
    w = torch.sqrt(aa.dot(aa) + 1e-16)
    if float(w.detach()) > 1e-9:
        v = (aa / w) * torch.sin(w / 2)
        return torch.cat([torch.cos(w / 2).reshape(1), v])
    return vec(1., 0., 0., 0.) + 0 * aa.sum()


def inv_trans(pos, position, rotation):                  # utils.py:50-54
    inv = torch.stack([rotation[0], -rotation[1], -rotation[2], -rotation[3]])
    inv = inv / torch.sqrt(inv.dot(inv))
    return qrot(inv, pos - position)


def length8(x):                                          # utils.py:4-6
    return torch.sqrt((x * x).sum(-1) + 1e-8)


def length14(x):                                         # primitives.py:13-16
    return torch.sqrt((x * x).sum(-1) + 1e-14)


# ---- tools (plb/engine/primitive/primitives.py, primive_base.py) ---------------------------------------------------------
# Primitive.default_config (primive_base.py:314-329) and Gripper.default_config (primitives.py:563-573)
PRIMITIVE_DEFAULTS = dict(init_pos=(0.3, 0.3, 0.3), init_rot=(1., 0., 0., 0.), lower_bound=(0., 0., 0.), upper_bound=(1., 1., 1.),
                          friction=0.9)
GRIPPER_DEFAULTS = dict(size=(0.03, 0.06, 0.03), minimal_gap=0.06, maximal_gap=1., init_gap=0.06)


class Tool:
    def __init__(self, cfg):
        cfg = {**PRIMITIVE_DEFAULTS, **(GRIPPER_DEFAULTS if cfg['shape'] == 'Gripper' else {}), **cfg}
        self.cfg = cfg
        self.shape = cfg['shape']
        a = cfg.get('action', {}) or {}
        self.action_dim = int(a.get('dim', 0))
        self.scale = torch.tensor(list(a.get('scale', ())) + [0.] * 8, dtype=DT)[:max(self.action_dim, 1)]
        self.friction = float(cfg.get('friction', 0.9))
        self.lo, self.hi = vec(*cfg.get('lower_bound', (0., 0., 0.))), vec(*cfg.get('upper_bound', (1., 1., 1.)))
        self.state_dim = 8 if self.shape in ('Gripper',) else 7
        if self.shape in ('Box', 'Gripper'):
            self.size = vec(*cfg['size'])
        if self.shape in ('RollingPinExt', 'Capsule'):
            self.h, self.r = float(cfg['h']), float(cfg['r'])
        if self.shape == 'Knife':
            self.h0, self.h1 = [float(v) for v in cfg['h']]
            self.size = vec(*cfg['size'])
            self.prot = vec(*cfg['prot'])
        if self.shape == 'Gripper':
            self.min_gap, self.max_gap = float(cfg['minimal_gap']), float(cfg['maximal_gap'])
        st = list(cfg['init_pos']) + list(cfg['init_rot'])
        if self.state_dim == 8:
            st.append(float(cfg['init_gap']))
        self.init_state = torch.tensor(st, dtype=DT)

    # local signed distances ------------------------------------------------------------------------------------------
    def box_sdf(self, p, size=None):                     # primitives.py:374-380
        q = p.abs() - (self.size if size is None else size)
        out = length14(torch.clamp(q, min=0.))
        return out + torch.clamp(q.max(-1).values, max=0.)

    def capsule_sdf(self, p):                            # primitives.py:54-61
        p2 = torch.stack([p[..., 0], p[..., 1] - torch.clamp(p[..., 1], -self.h / 2, self.h / 2), p[..., 2]], -1)
        return length14(p2) - self.r

    def prism_sdf(self, p0):                             # primitives.py:711-718
        inv = torch.stack([self.prot[0], -self.prot[1], -self.prot[2], -self.prot[3]])
        inv = inv / torch.sqrt(inv.dot(inv))
        p = qrot(inv, p0)
        q = p.abs()
        return torch.maximum(q[..., 2] - self.h1, torch.maximum(q[..., 0] * 0.866025 + p[..., 1] * 0.5, -p[..., 1]) - self.h0 * 0.5)

    def local_sdf(self, p):
        if self.shape == 'Box' or self.shape == 'Gripper':
            return self.box_sdf(p)
        if self.shape in ('RollingPinExt', 'Capsule'):
            return self.capsule_sdf(p)
        if self.shape == 'Knife':                        # primitives.py:769-773
            return torch.maximum(self.prism_sdf(p), self.box_sdf(p))
        raise NotImplementedError(self.shape)

    def local_normal(self, p):
        if self.shape in ('RollingPinExt', 'Capsule'):   # primitives.py:63-66: analytic
            p2 = torch.stack([p[..., 0], p[..., 1] - torch.clamp(p[..., 1], -self.h / 2, self.h / 2), p[..., 2]], -1)
            return p2 / length14(p2)[..., None]
        d = 1e-4                                         # primitives.py:383-393: central differences, then normalise
        cols = []
        for i in range(3):
            inc = torch.zeros(3, dtype=DT)
            inc[i] = d
            cols.append((1 / (2 * d)) * (self.local_sdf(p + inc) - self.local_sdf(p - inc)))
        n = torch.stack(cols, -1)
        return n / length14(n)[..., None]

    # contact of one rigid frame (primive_base.py:75-120); `eps14`: Gripper.collide2 uses the 1e-14 length (primitives.py:529)
    def collide_frame(self, pos0, rot0, pos1, rot1, gp, v, dt, softness, eps14=False):
        pl = inv_trans(gp, pos0, rot0)
        dist = self.local_sdf(pl)
        influence = torch.clamp(torch.exp(-dist * softness), max=1.)
        active = ((influence > 0.1) if softness > 0 else torch.zeros_like(dist, dtype=torch.bool)) | (dist <= 0)
        D = qrot(rot0, self.local_normal(pl))
        cv = (qrot(rot1, pl) + pos1 - gp) / dt           # collider_v: the relative position is the local point
        u = v - cv
        nc = (u * D).sum(-1)
        t = u - torch.clamp(nc, max=0.)[..., None] * D
        tn = (length14(t) if eps14 else length8(t))
        tf = t / tn[..., None] * torch.clamp(tn + nc * self.friction, min=0.)[..., None]
        flag = (nc < 0) & (torch.sqrt((t * t).sum(-1)) > 1e-30)
        t = torch.where(flag[..., None], tf, t)
        out = cv + u * (1 - influence)[..., None] + t * influence[..., None]
        return torch.where(active[..., None], out, v)

    def jaw(self, pos, rot, gap, sign):                  # Gripper.get_pos, primitives.py:471-473
        return pos + sign * qrot(rot, torch.stack([gap / 2, gap * 0, gap * 0]))

    def collide(self, s0, s1, gp, v, dt, softness):
        pos0, rot0, pos1, rot1 = s0[:3], s0[3:7], s1[:3], s1[3:7]
        if self.shape == 'Gripper':                      # primitives.py:507-536: the two jaws one after the other
            for sign in (-1., 1.):
                v = self.collide_frame(self.jaw(pos0, rot0, s0[7], sign), rot0, self.jaw(pos1, rot1, s1[7], sign), rot1, gp, v,
                                       dt, softness, eps14=True)
            return v
        return self.collide_frame(pos0, rot0, pos1, rot1, gp, v, dt, softness)

    # forward kinematics of one substep; a = the step's (clipped) action slice of this tool ------------------------------
    def fk(self, s, a, n_substeps):
        if self.action_dim == 0:
            v, w, gv = torch.zeros(3, dtype=DT), torch.zeros(3, dtype=DT), torch.zeros((), dtype=DT)
        else:                                            # set_velocity, primive_base.py:260-268 / primitives.py:462-469
            u = a * self.scale / n_substeps
            v = u[:3]
            w = u[3:6] if self.action_dim > 3 else torch.zeros(3, dtype=DT)
            gv = u[6] if self.action_dim > 6 else torch.zeros((), dtype=DT)
        pos, rot = s[:3], s[3:7]
        if self.shape == 'RollingPinExt':                # primitives.py:120-136
            dw, dth, dy = v[0], v[1], v[2]
            y_dir = qrot(rot, vec(0., -1., 0.))
            x_dir = torch.cross(vec(0., 1., 0.), y_dir, dim=-1) * (dw * 0.03 + w[0])
            x_dir = torch.stack([x_dir[0], dy, x_dir[2]])
            z = dw * 0
            rot1 = qmul(w2quat(torch.stack([z, -dth, z])), qmul(rot, w2quat(torch.stack([z, dw, z]))))
            pos1 = torch.maximum(torch.minimum(pos + x_dir, self.hi), self.lo)
            return torch.cat([pos1, rot1])
        pos1 = torch.maximum(torch.minimum(pos + v, self.hi), self.lo)
        if self.shape == 'Gripper':                      # primitives.py:456-460: right-multiply, clamp the gap
            gap1 = torch.clamp(s[7] - gv, self.min_gap, self.max_gap)
            return torch.cat([pos1, qmul(rot, w2quat(w)), gap1.reshape(1)])
        return torch.cat([pos1, qmul(w2quat(w), rot)])   # primive_base.py:152-156: left-multiply


# ---- backward_svd (mpm_simulator.py:131-156) as the backward of the SVD node -----------------------------------------------
class RefSVD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, F):
        U, s, Vh = torch.linalg.svd(F)
        V = Vh.transpose(-1, -2)
        # ti.svd contract: proper rotations, the sign on the last singular value
        du, dv = torch.linalg.det(U), torch.linalg.det(V)
        U = torch.cat([U[..., :2], U[..., 2:] * du[..., None, None]], -1)
        V = torch.cat([V[..., :2], V[..., 2:] * dv[..., None, None]], -1)
        s = torch.cat([s[..., :2], s[..., 2:] * (du * dv)[..., None]], -1)
        ctx.save_for_backward(U, s, V)
        return U, s, V

    @staticmethod
    def backward(ctx, gU, gs, gV):
        U, s, V = ctx.saved_tensors
        Ut, Vt = U.transpose(-1, -2), V.transpose(-1, -2)
        s2 = s * s
        d = s2[..., None, :] - s2[..., :, None]          # [i][j] = s_j^2 - s_i^2
        d = torch.where(d >= 0, torch.clamp(d, min=1e-6), torch.clamp(d, max=-1e-6))       # clamp(), :184-192
        Fm = (1.0 / d) * (1 - torch.eye(3, dtype=DT))
        S = torch.diag_embed(s)
        u_term = U @ ((Fm * (Ut @ gU - gU.transpose(-1, -2) @ U)) @ S) @ Vt
        v_term = U @ (S @ (Fm * (Vt @ gV - gV.transpose(-1, -2) @ V))) @ Vt
        return U @ torch.diag_embed(gs) @ Vt + u_term + v_term


class TorchMPM:
    """One env; state = (x, v, C, F) tensors [n, ...], tool states = list of [7 or 8] tensors."""

    def __init__(self, cfg, softness=666.):
        sim = cfg.SIMULATOR
        quality = sim.quality * sim.quality_multiplier * 0.5                 # mpm_simulator.py:19-21 (dim == 3)
        self.n = int(128 * quality)
        self.dx, self.inv_dx = 1 / self.n, float(self.n)
        self.dt = 0.5e-4 / quality
        self.p_vol = (self.dx * 0.5) ** 2
        self.p_mass = self.p_vol
        E, nu = sim.E, sim.nu
        self.mu, self.lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
        self.ys = sim.yield_stress
        self.substeps = int(2e-3 // self.dt)
        self.gravity = vec(*sim.gravity)
        self.ground_friction = float(sim.ground_friction)
        self.lower_bound = float(sim.lower_bound)
        self.softness = softness
        self.tools = [Tool(dict(p)) for p in cfg.PRIMITIVES]
        self.action_dims = [0]
        for t in self.tools:
            self.action_dims.append(self.action_dims[-1] + t.action_dim)
        n = self.n
        idx = torch.arange(n, dtype=DT)
        self.node_pos = torch.stack(torch.meshgrid(idx, idx, idx, indexing='ij'), -1).reshape(-1, 3) * self.dx
        self.node_I = torch.stack(torch.meshgrid(torch.arange(n), torch.arange(n), torch.arange(n), indexing='ij'), -1).reshape(-1, 3)

    def weights(self, x):                                                    # mpm_simulator.py:201-204
        xg = x * self.inv_dx
        base = (xg - 0.5).to(torch.int64)                                    # cast: truncation toward zero
        fx = xg - base.to(DT)
        w = torch.stack([0.5 * (1.5 - fx) ** 2, 0.75 - (fx - 1) ** 2, 0.5 * (fx - 0.5) ** 2], 0)   # [3, n, 3]
        return base, fx, w

    def substep(self, x, v, C, F, tool_states, action):
        """One substep (mpm_simulator.py:307-323).  action: the env step's clipped action [A]; returns the new state."""
        n, dt, dx, inv_dx, p_mass = self.n, self.dt, self.dx, self.inv_dx, self.p_mass
        nxt = [t.fk(s, action[self.action_dims[i]:self.action_dims[i + 1]], self.substeps) for i, (t, s) in
               enumerate(zip(self.tools, tool_states))]
        I3 = torch.eye(3, dtype=DT)
        Ftmp = (I3 + dt * C) @ F                                             # compute_F_tmp :121-124
        U, sig, V = RefSVD.apply(Ftmp)
        # compute_von_mises :165-182
        sc = torch.clamp(sig, min=0.05)
        eps = torch.log(sc)
        eh = eps - eps.sum(-1, keepdim=True) / 3
        ehn = torch.sqrt((eh * eh).sum(-1) + 1e-8)
        dg = ehn - self.ys / (2 * self.mu)
        e_new = torch.exp(eps - (dg / ehn)[..., None] * eh)
        F_yield = U @ torch.diag_embed(e_new) @ V.transpose(-1, -2)
        newF = torch.where((dg > 0)[..., None, None], F_yield, Ftmp)
        J = torch.linalg.det(newF)
        r = U @ V.transpose(-1, -2)
        stress = 2 * self.mu * (newF - r) @ newF.transpose(-1, -2) + (self.lam * J * (J - 1))[..., None, None] * I3
        stress = (-dt * self.p_vol * 4 * inv_dx * inv_dx) * stress
        affine = stress + p_mass * C
        base, fx, w = self.weights(x)
        G = n * n * n
        gv_in = torch.zeros((G, 3), dtype=DT)
        gm = torch.zeros((G,), dtype=DT)
        for i in range(3):                                                   # p2g :216-225
            for j in range(3):
                for l in range(3):
                    off = vec(float(i), float(j), float(l))
                    dpos = (off - fx) * dx
                    wt = w[i, :, 0] * w[j, :, 1] * w[l, :, 2]
                    node = ((base[:, 0] + i) * n + base[:, 1] + j) * n + base[:, 2] + l
                    gv_in = gv_in.index_add(0, node, wt[:, None] * (p_mass * v + (affine @ dpos[..., None])[..., 0]))
                    gm = gm.index_add(0, node, wt * p_mass)
        # grid_op :230-262 on the occupied nodes
        occ = torch.nonzero(gm > 1e-12)[:, 0]
        vo = gv_in[occ] / gm[occ][:, None] + dt * self.gravity * 30
        gp, Iocc = self.node_pos[occ], self.node_I[occ]
        for t, s0, s1 in zip(self.tools, tool_states, nxt):
            vo = t.collide(s0, s1, gp, vo, dt, self.softness)
        bound = 3
        for d in range(3):
            low = (Iocc[:, d] < bound) & (vo[:, d] < 0)
            if d != 1 or self.ground_friction == 0:
                comp = torch.where(low, torch.zeros_like(vo[:, d]), vo[:, d])
                vo = torch.cat([vo[:, :d], comp[:, None], vo[:, d + 1:]], 1)
            elif self.ground_friction < 10:                                  # Coulomb branch :246-256
                nrm = vec(0., 1., 0.)
                lin = vo[:, 1] + 1e-30
                vit = vo - lin[:, None] * nrm - Iocc.to(DT) * 1e-30
                lit = torch.sqrt((vit * vit).sum(-1) + 1e-8)
                vfr = torch.clamp(1 + self.ground_friction * lin / lit, min=0.)[:, None] * (vit + Iocc.to(DT) * 1e-30)
                vfr = torch.cat([vfr[:, :1], torch.zeros_like(vfr[:, 1:2]), vfr[:, 2:]], 1)
                vo = torch.where(low[:, None], vfr, vo)
            else:
                vo = torch.where(low[:, None], torch.zeros_like(vo), vo)
            high = (Iocc[:, d] > n - bound) & (vo[:, d] > 0)
            comp = torch.where(high, torch.zeros_like(vo[:, d]), vo[:, d])
            vo = torch.cat([vo[:, :d], comp[:, None], vo[:, d + 1:]], 1)
        gv_out = torch.zeros((G, 3), dtype=DT).index_add(0, occ, vo)
        # g2p :264-283
        new_v = torch.zeros_like(v)
        new_C = torch.zeros_like(C)
        for i in range(3):
            for j in range(3):
                for l in range(3):
                    off = vec(float(i), float(j), float(l))
                    dpos = off - fx
                    wt = w[i, :, 0] * w[j, :, 1] * w[l, :, 2]
                    node = ((base[:, 0] + i) * n + base[:, 1] + j) * n + base[:, 2] + l
                    g = gv_out[node]
                    new_v = new_v + wt[:, None] * g
                    new_C = new_C + 4 * inv_dx * wt[:, None, None] * (g[:, :, None] * dpos[:, None, :])
        new_x = torch.clamp(x + dt * new_v, self.lower_bound * dx, 1 - 3 * dx)
        return new_x, new_v, new_C, newF, nxt

    def step(self, x, v, C, F, tool_states, action):
        a = torch.clamp(action, -1, 1)                                       # primitives.py:864
        for _ in range(self.substeps):
            x, v, C, F, tool_states = self.substep(x, v, C, F, tool_states, a)
        return x, v, C, F, tool_states
