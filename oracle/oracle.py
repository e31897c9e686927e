"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/mpm_oracle.h).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package ``diffskill_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_TOOLS, MAX_PAIRS = 8, 8


class ToolCfg(C.Structure):
    _fields_ = [('type', C.c_int32), ('action_dim', C.c_int32), ('action_scale', C.c_double * 8),
                ('friction', C.c_double), ('softness', C.c_double),
                ('lower_bound', C.c_double * 3), ('upper_bound', C.c_double * 3), ('size', C.c_double * 3),
                ('h', C.c_double), ('r', C.c_double), ('radius', C.c_double), ('prism_h', C.c_double * 2),
                ('prot', C.c_double * 4), ('minimal_gap', C.c_double), ('maximal_gap', C.c_double)]


class Config(C.Structure):
    _fields_ = [('n_grid', C.c_int32), ('substeps', C.c_int32), ('max_frames', C.c_int32),
                ('max_particles', C.c_int32),
                ('dt', C.c_double), ('dx', C.c_double), ('inv_dx', C.c_double), ('p_vol', C.c_double),
                ('p_mass', C.c_double), ('mu', C.c_double), ('lam', C.c_double), ('yield_stress', C.c_double),
                ('gravity', C.c_double * 3), ('ground_friction', C.c_double), ('lower_bound', C.c_double),
                ('n_tools', C.c_int32), ('tools', ToolCfg * MAX_TOOLS),
                ('n_pairs', C.c_int32), ('pairs', (C.c_int32 * 2) * MAX_PAIRS)]


def build(force=False):
    so = os.path.join(_HERE, 'libmpm_oracle.so')
    srcs = [os.path.join(_HERE, f) for f in ('mpm_oracle.cpp', 'mpm_oracle.h', 'ad.hpp')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'libmpm_oracle.so')
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_tool_sdf.restype = C.c_double
        _LIB = L
    return _LIB


def set_svd_algorithm(alg):
    """0: one-sided Jacobi (default), 1: eigen-decomposition of A^T A + Givens QR (process-wide)."""
    lib().orc_set_svd_algorithm(int(alg))


def set_scatter_noise(seed, ulps=1.0):
    """Process-wide: perturb every grid sum of p2g / g2p.grad by -ulps/0/+ulps units in the last place (hash of seed,
    frame, node) -- the run-to-run noise of the reference's unordered float atomics.  The spread of a result over a few
    seeds is the reference formulation's own reproducibility floor on a scene.  seed 0 switches it off."""
    lib().orc_set_scatter_noise(int(seed), C.c_double(ulps))


def set_pose_adjoint_atomics(seed):
    """Process-wide: grid_op.grad adds every node's tool-pose adjoint terms one by one in the simulation precision, in a
    seeded random node order (what the reference's float atomics do); 0 = order-independent double sum (default)."""
    lib().orc_set_pose_adjoint_atomics(int(seed))


def set_fast_math_noise(amplitude, salt=1):
    """Process-wide: log gets a pseudo-random absolute error in [-a, a], exp a relative one in [-a/2, a/2] (the reference
    runs ti.init(fast_math=True): hardware log/exp, |err(log)| <= 2^-21.4).  amplitude 0 switches it off."""
    lib().orc_set_fast_math_noise(C.c_double(amplitude), int(salt))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def make_config(scene, max_frames, max_particles, softness=666.):
    c = Config()
    c.n_grid, c.substeps, c.max_frames, c.max_particles = scene.n_grid, scene.substeps, max_frames, max_particles
    c.dt, c.dx, c.inv_dx, c.p_vol, c.p_mass = scene.dt, scene.dx, scene.inv_dx, scene.p_vol, scene.p_mass
    c.mu, c.lam, c.yield_stress = scene.mu, scene.lam, scene.yield_stress
    c.gravity[:] = scene.gravity
    c.ground_friction, c.lower_bound = scene.ground_friction, scene.lower_bound
    c.n_tools = len(scene.tools)
    for i, t in enumerate(scene.tools):
        tc = c.tools[i]
        tc.type, tc.action_dim = t.type_id, t.action_dim
        for j, s in enumerate(t.action_scale[:8]):
            tc.action_scale[j] = s
        tc.friction, tc.softness = t.friction, softness
        tc.lower_bound[:] = t.lower_bound
        tc.upper_bound[:] = t.upper_bound
        tc.size[:] = t.size
        tc.h, tc.r, tc.radius = t.h, t.r, t.radius
        tc.prism_h[:] = t.prism_h
        tc.prot[:] = t.prot
        tc.minimal_gap, tc.maximal_gap = t.minimal_gap, t.maximal_gap
    c.n_pairs = len(scene.pairs)
    for k, (i, j) in enumerate(scene.pairs):
        c.pairs[k][0], c.pairs[k][1] = i, j
    return c


class Oracle:
    """Frame-taped simulator with the reference's method names (MPMSimulator / Primitives / GradModel)."""

    def __init__(self, scene, n_particles, max_frames, f64=False, softness=666., threads=None):
        self.L = lib()
        self.scene = scene
        self.f64 = f64
        self.substeps = scene.substeps
        self.cfg = make_config(scene, max_frames, n_particles, softness)
        self.h = C.c_void_p(self.L.orc_create(C.byref(self.cfg), int(f64)))
        self.n = n_particles
        self.G = scene.n_grid ** 3
        self.K = len(scene.tools)
        if threads is not None:
            self.L.orc_set_threads(int(threads))
        self.L.orc_initialize(self.h, n_particles)
        if scene.pairs:
            _, p = _d(scene.rand_num())
            self.L.orc_set_rand_num(self.h, p)
        for i, t in enumerate(scene.tools):
            self.set_tool_state(0, i, t.init_state)
        self.cur = 0

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # ---- state -------------------------------------------------------------
    def reset(self, x):
        n = len(x)
        self.set_frame(0, x, np.zeros((n, 3)), np.tile(np.eye(3), (n, 1, 1)), np.zeros((n, 3, 3)))

    def set_frame(self, f, x, v, F, Cm):
        self.n = len(x)
        (_, a), (_, b), (_, c), (_, d) = _d(x), _d(v), _d(F), _d(Cm)
        self.L.orc_set_frame(self.h, f, self.n, a, b, c, d)

    def get_frame(self, f):
        n = self.n
        x, v, F, Cm = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
        self.L.orc_get_frame(self.h, f, _d(x)[1], _d(v)[1], _d(F)[1], _d(Cm)[1])
        return x, v, F, Cm

    def set_tool_state(self, f, i, st):
        s = np.zeros(8)
        s[:len(st)] = st
        self.L.orc_set_tool_state(self.h, f, i, _d(s)[1])

    def get_tool_state(self, f, i):
        s = np.zeros(8)
        self.L.orc_get_tool_state(self.h, f, i, _d(s)[1])
        return s

    def get_tool_states(self, f):
        return np.stack([self.get_tool_state(f, i) for i in range(self.K)])

    def set_tool_param(self, i, which, value):
        self.L.orc_set_tool_param(self.h, i, which, C.c_double(value))

    def set_material(self, mu=None, lam=None, ys=None):
        p = [None if a is None else _d(np.broadcast_to(a, (self.n,)))[1] for a in (mu, lam, ys)]
        self.L.orc_set_material(self.h, *p)

    def copyframe(self, src, dst):
        self.L.orc_copyframe(self.h, src, dst)

    # ---- stepping ------------------------------------------------------------
    def set_action(self, s, action, n_substeps=None):
        self.L.orc_set_action(self.h, s, n_substeps or self.substeps, _d(action)[1])

    def substep(self, f):
        self.L.orc_substep(self.h, f)

    def substep_grad(self, f):
        self.L.orc_substep_grad(self.h, f)

    def forward_step(self, s, action):       # function.py:166-175
        self.set_action(s, action)
        for f in range(s * self.substeps, (s + 1) * self.substeps):
            self.substep(f)

    def backward_step(self, s):              # function.py:177-190
        for f in range((s + 1) * self.substeps - 1, s * self.substeps - 1, -1):
            self.substep_grad(f)
        self.L.orc_set_velocity_grad(self.h, s, self.substeps)
        return self.get_action_grad(s)

    def step_copy(self, action):             # mpm_simulator.py:440-451 with is_copy=True
        self.set_action(0, action)
        for f in range(self.substeps):
            self.substep(f)
        self.copyframe(self.substeps, 0)

    def get_action_grad(self, s):
        out = np.zeros(self.scene.action_dim)
        self.L.orc_get_action_grad(self.h, s, _d(out)[1])
        return out

    # ---- adjoints ------------------------------------------------------------
    def zero_grad(self):
        self.L.orc_zero_grad(self.h)

    def get_frame_grad(self, f):
        n = self.n
        x, v, F, Cm = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3, 3))
        self.L.orc_get_frame_grad(self.h, f, _d(x)[1], _d(v)[1], _d(F)[1], _d(Cm)[1])
        return x, v, F, Cm

    def add_frame_grad(self, f, gx=None, gv=None, gF=None, gC=None):
        p = [None if a is None else _d(a)[1] for a in (gx, gv, gF, gC)]
        self.L.orc_add_frame_grad(self.h, f, *p)

    def scale_frame_grad(self, f, alpha):
        self.L.orc_scale_frame_grad(self.h, f, C.c_double(alpha))

    def get_tool_grad(self, f, i):
        g = np.zeros(8)
        self.L.orc_get_tool_grad(self.h, f, i, _d(g)[1])
        return g

    def get_tool_grads(self, f):
        return np.stack([self.get_tool_grad(f, i) for i in range(self.K)])

    def add_tool_grad(self, f, i, g8):
        self.L.orc_add_tool_grad(self.h, f, i, _d(g8)[1])

    def get_tool_vel_grad(self, f, i):
        g = np.zeros(7)
        self.L.orc_get_tool_vel_grad(self.h, f, i, _d(g)[1])
        return g

    # ---- inspection ------------------------------------------------------------
    def get_grid(self):
        n = self.scene.n_grid
        a, b, m = np.zeros((n, n, n, 3)), np.zeros((n, n, n, 3)), np.zeros((n, n, n))
        self.L.orc_get_grid(self.h, _d(a)[1], _d(b)[1], _d(m)[1])
        return a, b, m

    def get_grid_grad(self):
        n = self.scene.n_grid
        a, b, m = np.zeros((n, n, n, 3)), np.zeros((n, n, n, 3)), np.zeros((n, n, n))
        self.L.orc_get_grid_grad(self.h, _d(a)[1], _d(b)[1], _d(m)[1])
        return a, b, m

    def get_svd(self):
        n = self.n
        Ft, U, s, V = np.zeros((n, 3, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3)), np.zeros((n, 3, 3))
        self.L.orc_get_svd(self.h, _d(Ft)[1], _d(U)[1], _d(s)[1], _d(V)[1])
        return Ft, U, s, V

    def cell_index(self, f):
        base = np.zeros((self.n, 3), np.int32)
        key = np.zeros(self.n, np.int32)
        self.L.orc_cell_index(self.h, f, base.ctypes.data_as(C.POINTER(C.c_int32)),
                              key.ctypes.data_as(C.POINTER(C.c_int32)))
        return base, key

    def occupancy(self, f):
        n = self.scene.n_grid
        occ = np.zeros((n, n, n), np.uint8)
        self.L.orc_occupancy(self.h, f, occ.ctypes.data_as(C.POINTER(C.c_uint8)))
        return occ

    def collision_idx(self, f):
        idx = np.zeros(max(1, len(self.scene.pairs)), np.int32)
        self.L.orc_get_collision_idx(self.h, f, idx.ctypes.data_as(C.POINTER(C.c_int32)))
        return idx[:len(self.scene.pairs)]

    def compute_min_dist(self, f):
        nc = self.L.orc_min_dist_cols(self.h)
        out = np.zeros((self.n, nc))
        self.L.orc_compute_min_dist(self.h, f, _d(out)[1])
        return out

    def compute_min_dist_grad(self, f, gin):
        self.L.orc_compute_min_dist_grad(self.h, f, _d(gin)[1])

    def compute_grid_m(self, f):
        n = self.scene.n_grid
        out = np.zeros((n, n, n))
        self.L.orc_compute_grid_m(self.h, f, _d(out)[1])
        return out

    def compute_grid_m_grad(self, f, gin):
        self.L.orc_compute_grid_m_grad(self.h, f, _d(gin)[1])

    # ---- probes ------------------------------------------------------------------
    def tool_sdf(self, i, f, p):
        return self.L.orc_tool_sdf(self.h, i, f, _d(p)[1])

    def tool_normal(self, i, f, p):
        n = np.zeros(3)
        self.L.orc_tool_normal(self.h, i, f, _d(p)[1], _d(n)[1])
        return n

    def tool_probe_grad(self, i, f, what, p, v, gout):
        """Tape-AD adjoint of tool_sdf ('sdf'), tool_normal ('normal') or tool_collide ('collide') of tool i at poses
        f, f+1 -> (g_p[3], g_v_in[3], g_pose_f[8], g_pose_f1[8])."""
        out = np.zeros(22)
        g = np.zeros(3)
        g[:np.size(gout)] = np.ravel(gout)
        pa, pp = _d(p)
        va, vp = _d(v)
        ga, gp = _d(g)
        self.L.orc_tool_probe_grad(self.h, int(i), int(f), {'sdf': 0, 'normal': 1, 'collide': 2}[what], pp, vp, gp,
                                   out.ctypes.data_as(C.POINTER(C.c_double)))
        return out[0:3], out[3:6], out[6:14], out[14:22]

    def tool_probe_fk(self, i, state8, vel7, gnext8=None):
        """forward_kinematics of tool i from an explicit state -> next8 (and (g_state[8], g_vel[7]) with gnext8)."""
        sa, sp = _d(state8)
        va, vp = _d(vel7)
        nxt, g = np.zeros(8), np.zeros(15)
        if gnext8 is None:
            self.L.orc_tool_probe_fk(self.h, int(i), sp, vp, nxt.ctypes.data_as(C.POINTER(C.c_double)), None, None)
            return nxt
        ga, gp = _d(gnext8)
        self.L.orc_tool_probe_fk(self.h, int(i), sp, vp, nxt.ctypes.data_as(C.POINTER(C.c_double)), gp,
                                 g.ctypes.data_as(C.POINTER(C.c_double)))
        return nxt, g[:8], g[8:]

    def tool_collide(self, i, f, p, v):
        o = np.zeros(3)
        self.L.orc_tool_collide(self.h, i, f, _d(p)[1], _d(v)[1], _d(o)[1])
        return o


def svd3(F, f64=False):
    U, s, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    lib().orc_svd3(int(f64), _d(F)[1], _d(U)[1], _d(s)[1], _d(V)[1])
    return U, s, V
