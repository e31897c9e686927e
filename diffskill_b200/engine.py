"""ctypes binding of libdiffskill_mpm.so -- the only compute path of this package.

There is no CPU fallback: if the CUDA library is missing or no sm_100 device is
visible, constructing an :class:`Engine` raises.  Arrays cross the boundary as
raw pointers (numpy host arrays, or torch CUDA tensors via ``data_ptr()``).
"""
import ctypes as C
import os

import numpy as np

from .scene import NUM_COLLISION_POINTS, TOOL_SPHERE, SceneSpec

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdiffskill_mpm.so')
if os.environ.get('DSK_LIB') == 'timeline':   # profiling build with in-graph kernel timestamps (build.py --timeline)
    LIB_PATH = os.path.join(_HERE, 'libdiffskill_mpm_tl.so')
if os.environ.get('DSK_LIB') == 'precise':    # diagnostic build without fast-math log/exp/div/rsqrt (build.py --precise)
    LIB_PATH = os.path.join(_HERE, 'libdiffskill_mpm_pm.so')
MAX_TOOLS, MAX_PAIRS, ABI_VERSION = 8, 8, 1

PARAM_FRICTION, PARAM_SOFTNESS, PARAM_LOWER, PARAM_UPPER = 0, 1, 2, 5


class ToolDesc(C.Structure):
    _fields_ = [('type', C.c_int32), ('action_dim', C.c_int32), ('action_scale', C.c_double * 8),
                ('friction', C.c_double), ('softness', C.c_double),
                ('lower_bound', C.c_double * 3), ('upper_bound', C.c_double * 3), ('size', C.c_double * 3),
                ('h', C.c_double), ('r', C.c_double), ('prism_h', C.c_double * 2), ('prot', C.c_double * 4),
                ('minimal_gap', C.c_double), ('maximal_gap', C.c_double)]


class Config(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('device', C.c_int32), ('n_envs', C.c_int32),
                ('particle_capacity', C.c_int32), ('n_grid', C.c_int32), ('substeps', C.c_int32),
                ('max_steps', C.c_int32), ('step_slots', C.c_int32), ('sort_particles', C.c_int32),
                ('grid_tape_mib', C.c_int32),
                ('dt', C.c_double), ('dx', C.c_double), ('inv_dx', C.c_double), ('p_vol', C.c_double),
                ('p_mass', C.c_double), ('mu', C.c_double), ('lam', C.c_double), ('yield_stress', C.c_double),
                ('gravity', C.c_double * 3), ('ground_friction', C.c_double), ('lower_bound', C.c_double),
                ('n_tools', C.c_int32), ('n_pairs', C.c_int32), ('tools', ToolDesc * MAX_TOOLS),
                ('pairs', (C.c_int32 * 2) * MAX_PAIRS)]


class EngineError(RuntimeError):
    pass


_LIB = None

# every symbol include/diffskill_mpm.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    'dsk_last_error', 'dsk_abi_version', 'dsk_sizeof_config', 'dsk_sizeof_tool_desc', 'dsk_create', 'dsk_destroy', 'dsk_set_stream', 'dsk_synchronize',
    'dsk_set_rand_num', 'dsk_set_particles', 'dsk_get_particles', 'dsk_get_n_particles', 'dsk_set_tool_state',
    'dsk_get_tool_state', 'dsk_copy_step', 'dsk_set_material', 'dsk_set_tool_param', 'dsk_get_tool_param',
    'dsk_set_gravity', 'dsk_set_action', 'dsk_forward_step', 'dsk_backward_step', 'dsk_substep', 'dsk_substep_grad',
    'dsk_zero_grad', 'dsk_add_particle_grad', 'dsk_add_tool_grad', 'dsk_get_particle_grad', 'dsk_get_tool_grad',
    'dsk_scale_grad', 'dsk_get_action_grad', 'dsk_get_action_grads', 'dsk_get_obs', 'dsk_min_dist_cols', 'dsk_compute_min_dist',
    'dsk_compute_min_dist_grad', 'dsk_compute_grid_m', 'dsk_compute_grid_m_grad', 'dsk_debug_cell_index',
    'dsk_debug_sort_order', 'dsk_debug_grid', 'dsk_debug_grid_grad', 'dsk_debug_frame', 'dsk_debug_tool_frame',
    'dsk_debug_tool_frame_grad', 'dsk_debug_svd', 'dsk_launch_count', 'dsk_memory_bytes', 'dsk_set_graphs',
    'dsk_profile_enable', 'dsk_kernel_class_count', 'dsk_kernel_class_name', 'dsk_profile_report',
    'dsk_launch_counts', 'dsk_loss_reset', 'dsk_loss_add_l2', 'dsk_loss_get',
    'dsk_timeline_enable', 'dsk_timeline_reset', 'dsk_timeline_read',
    'dsk_set_actions', 'dsk_forward_steps', 'dsk_backward_steps', 'dsk_loss_add_l2_steps',
]


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: build it with `python -m diffskill_b200.build` "
                              "(or __graft_entry__.build()). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.dsk_last_error.restype = C.c_char_p
        L.dsk_kernel_class_name.restype = C.c_char_p
        _LIB = L
    return _LIB


def _ptr(a):
    """(pointer, on_device) of a numpy array / torch tensor / None."""
    if a is None:
        return None, 0
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
        return C.c_void_p(a.ctypes.data), 0
    # torch tensor (duck-typed so that importing this module does not need torch)
    assert a.is_contiguous() and str(a.dtype) == 'torch.float32', "need contiguous float32 tensors"
    return C.c_void_p(a.data_ptr()), int(a.is_cuda)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def make_config(scene: SceneSpec, n_envs, capacity, max_steps, step_slots, sort, softness, device, grid_tape_mib=0):
    c = Config()
    c.abi_version, c.device, c.n_envs, c.particle_capacity = ABI_VERSION, device, n_envs, capacity
    c.n_grid, c.substeps, c.max_steps, c.step_slots, c.sort_particles = \
        scene.n_grid, scene.substeps, max_steps, step_slots, int(sort)
    c.grid_tape_mib = int(grid_tape_mib)
    c.dt, c.dx, c.inv_dx, c.p_vol, c.p_mass = scene.dt, scene.dx, scene.inv_dx, scene.p_vol, scene.p_mass
    c.mu, c.lam, c.yield_stress = scene.mu, scene.lam, scene.yield_stress
    c.gravity[:] = scene.gravity
    c.ground_friction, c.lower_bound = scene.ground_friction, scene.lower_bound
    c.n_tools, c.n_pairs = len(scene.tools), len(scene.pairs)
    for i, t in enumerate(scene.tools):
        d = c.tools[i]
        d.type, d.action_dim = t.type_id, t.action_dim
        for j, s in enumerate(t.action_scale[:8]):
            d.action_scale[j] = s
        d.friction, d.softness = t.friction, softness
        d.lower_bound[:] = t.lower_bound
        d.upper_bound[:] = t.upper_bound
        d.size[:] = t.size
        d.h, d.r = t.h, (t.radius if t.type_id == TOOL_SPHERE else t.r)
        d.prism_h[:] = t.prism_h
        d.prot[:] = t.prot
        d.minimal_gap, d.maximal_gap = t.minimal_gap, t.maximal_gap
    for k, (i, j) in enumerate(scene.pairs):
        c.pairs[k][0], c.pairs[k][1] = i, j
    return c


class Engine:
    """One batched simulation engine (B envs sharing a scene).  Thin, 1:1 over the C ABI."""

    def __init__(self, scene: SceneSpec, n_envs=1, capacity=None, max_steps=64, step_slots=1, sort=True,
                 softness=666., device=0, grid_tape_mib=256):
        self.L = load_library()
        self.scene = scene
        self.B = int(n_envs)
        self.capacity = int(capacity or scene.particle_capacity)
        self.H = int(max_steps)
        self.S = scene.substeps
        self.K = len(scene.tools)
        self.A = scene.action_dim
        self.n_grid = scene.n_grid
        self.device = device
        self.cfg = make_config(scene, self.B, self.capacity, self.H, step_slots, sort, softness, device, grid_tape_mib)
        h = C.c_void_p()
        self._ck(self.L.dsk_create(C.byref(self.cfg), C.byref(h)))
        self.h = h
        if scene.pairs:
            rn = np.ascontiguousarray(scene.rand_num(), dtype=np.float64)
            self._ck(self.L.dsk_set_rand_num(self.h, rn.ctypes.data_as(C.POINTER(C.c_double))))
        for b in range(self.B):
            for i, t in enumerate(scene.tools):
                self.set_tool_state(0, b, i, t.init_state)
        nc = C.c_int()
        self._ck(self.L.dsk_min_dist_cols(self.h, C.byref(nc)))
        self.ncols = nc.value

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.L.dsk_last_error().decode())

    def close(self):
        if getattr(self, 'h', None):
            self.L.dsk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._ck(self.L.dsk_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.L.dsk_synchronize(self.h))

    # ---- state ---------------------------------------------------------------------------------------
    def set_particles(self, step, env, x, v=None, F=None, Cm=None):
        if not isinstance(x, np.ndarray) and hasattr(x, 'data_ptr'):
            n = x.shape[0]
            (px, dev), (pv, _), (pF, _), (pC, _) = _ptr(x), _ptr(v), _ptr(F), _ptr(Cm)
        else:
            x = _f32(x)
            n = len(x)
            v = _f32(np.zeros((n, 3)) if v is None else v)
            F = _f32(np.tile(np.eye(3), (n, 1, 1)) if F is None else F)
            Cm = _f32(np.zeros((n, 3, 3)) if Cm is None else Cm)
            (px, dev), (pv, _), (pF, _), (pC, _) = _ptr(x), _ptr(v), _ptr(F), _ptr(Cm)
        self._ck(self.L.dsk_set_particles(self.h, step, env, n, px, pv, pF, pC, dev))

    def n_particles(self, env=0):
        n = C.c_int()
        self._ck(self.L.dsk_get_n_particles(self.h, env, C.byref(n)))
        return n.value

    def get_particles(self, step, env=0, fields='xvFC'):
        n = self.n_particles(env)
        out = {}
        shapes = {'x': (n, 3), 'v': (n, 3), 'F': (n, 3, 3), 'C': (n, 3, 3)}
        for f in 'xvFC':
            out[f] = np.zeros(shapes[f], np.float32) if f in fields else None
        p = [_ptr(out[f])[0] for f in 'xvFC']
        self._ck(self.L.dsk_get_particles(self.h, step, env, p[0], p[1], p[2], p[3], 0))
        return tuple(out[f] for f in fields)

    def set_tool_state(self, step, env, tool, state):
        s = np.zeros(8, np.float32)
        s[:len(state)] = state
        self._ck(self.L.dsk_set_tool_state(self.h, step, env, tool, _ptr(s)[0]))

    def get_tool_state(self, step, env, tool):
        s = np.zeros(8, np.float32)
        self._ck(self.L.dsk_get_tool_state(self.h, step, env, tool, _ptr(s)[0]))
        return s

    def get_tool_states(self, step, env=0):
        return np.stack([self.get_tool_state(step, env, i) for i in range(self.K)]) if self.K else np.zeros((0, 8))

    def copy_step(self, src, dst):
        self._ck(self.L.dsk_copy_step(self.h, src, dst))

    def set_material(self, env, mu=None, lam=None, yield_stress=None):
        n = self.n_particles(env)
        arrs = [None if a is None else _f32(np.broadcast_to(a, (n,))) for a in (mu, lam, yield_stress)]
        self._ck(self.L.dsk_set_material(self.h, env, *[_ptr(a)[0] for a in arrs]))

    def set_tool_param(self, tool, which, value):
        self._ck(self.L.dsk_set_tool_param(self.h, tool, which, C.c_double(value)))

    def get_tool_param(self, tool, which):
        v = C.c_double()
        self._ck(self.L.dsk_get_tool_param(self.h, tool, which, C.byref(v)))
        return v.value

    def set_softness(self, softness):
        for i in range(self.K):
            self.set_tool_param(i, PARAM_SOFTNESS, softness)

    def set_gravity(self, g):
        a = (C.c_double * 3)(*g)
        self._ck(self.L.dsk_set_gravity(self.h, a))

    # ---- stepping ------------------------------------------------------------------------------------
    def set_action(self, step, actions):
        if isinstance(actions, np.ndarray) or not hasattr(actions, 'data_ptr'):
            actions = _f32(np.asarray(actions).reshape(self.B, self.A))
        p, dev = _ptr(actions)
        self._ck(self.L.dsk_set_action(self.h, step, p, dev))

    def forward_step(self, src, dst=None, action_step=None):
        dst = src + 1 if dst is None else dst
        action_step = src if action_step is None else action_step
        self._ck(self.L.dsk_forward_step(self.h, src, dst, action_step))

    def backward_step(self, step):
        self._ck(self.L.dsk_backward_step(self.h, step))

    def substep(self, f):
        self._ck(self.L.dsk_substep(self.h, f))

    def substep_grad(self, f):
        self._ck(self.L.dsk_substep_grad(self.h, f))

    # ---- adjoints ------------------------------------------------------------------------------------
    def zero_grad(self):
        self._ck(self.L.dsk_zero_grad(self.h))

    def add_particle_grad(self, step, gx=None, gv=None, gF=None, gC=None):
        """gx,gv: [B,capacity,3]; gF,gC: [B,capacity,3,3] (numpy host or torch CUDA)."""
        ps = [_ptr(_f32(a) if isinstance(a, np.ndarray) else a) for a in (gx, gv, gF, gC)]
        dev = max(d for _, d in ps)
        assert all(p is None or d == dev for p, d in ps), "mix of host and device gradients"
        self._keep = (gx, gv, gF, gC)
        self._ck(self.L.dsk_add_particle_grad(self.h, step, ps[0][0], ps[1][0], ps[2][0], ps[3][0], dev))

    def pad_particles(self, a, env_n=None):
        """[n, ...] -> [1?, capacity, ...] helper for single-env host gradients."""
        a = np.asarray(a, dtype=np.float32)
        out = np.zeros((self.capacity,) + a.shape[1:], np.float32)
        out[:len(a)] = a
        return out

    def add_tool_grad(self, step, g):
        if isinstance(g, np.ndarray) or not hasattr(g, 'data_ptr'):
            g = _f32(np.asarray(g).reshape(self.B, self.K, 8))
        p, dev = _ptr(g)
        self._ck(self.L.dsk_add_tool_grad(self.h, step, p, dev))

    def get_particle_grad(self, step, env=0):
        n = self.n_particles(env)
        gx, gv = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        gF, gC = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3, 3), np.float32)
        self._ck(self.L.dsk_get_particle_grad(self.h, step, env, _ptr(gx)[0], _ptr(gv)[0], _ptr(gF)[0], _ptr(gC)[0], 0))
        return gx, gv, gF, gC

    def get_tool_grad(self, step, env, tool):
        g = np.zeros(8, np.float32)
        self._ck(self.L.dsk_get_tool_grad(self.h, step, env, tool, _ptr(g)[0]))
        return g

    def get_tool_grads(self, step, env=0):
        return np.stack([self.get_tool_grad(step, env, i) for i in range(self.K)])

    def scale_grad(self, step, alpha):
        self._ck(self.L.dsk_scale_grad(self.h, step, C.c_double(alpha)))

    def get_action_grad(self, step, out=None):
        if out is None:
            out = np.zeros((self.B, self.A), np.float32)
        p, dev = _ptr(out)
        self._ck(self.L.dsk_get_action_grad(self.h, step, p, dev))
        return out

    def get_action_grads(self, step0=0, nsteps=None, out=None):
        nsteps = self.H - step0 if nsteps is None else nsteps
        if out is None:
            out = np.zeros((nsteps, self.B, self.A), np.float32)
        p, dev = _ptr(out)
        self._ck(self.L.dsk_get_action_grads(self.h, step0, nsteps, p, dev))
        return out

    # ---- observations --------------------------------------------------------------------------------
    def get_obs(self, step, xv=None, tools=None):
        if xv is None:
            xv = np.zeros((self.B, self.capacity, 6), np.float32)
        if tools is None:
            tools = np.zeros((self.B, self.K, 8), np.float32) if isinstance(xv, np.ndarray) else None
        (p, dev), (q, _) = _ptr(xv), _ptr(tools)
        self._ck(self.L.dsk_get_obs(self.h, step, p, q, dev))
        return xv, tools

    def compute_min_dist(self, step, out=None):
        if out is None:
            out = np.zeros((self.B, self.capacity, self.ncols), np.float32)
        p, dev = _ptr(out)
        self._ck(self.L.dsk_compute_min_dist(self.h, step, p, dev))
        return out

    def compute_min_dist_grad(self, step, g):
        p, dev = _ptr(_f32(g) if isinstance(g, np.ndarray) else g)
        self._ck(self.L.dsk_compute_min_dist_grad(self.h, step, p, dev))

    def compute_grid_m(self, step, out=None):
        n = self.n_grid
        if out is None:
            out = np.zeros((self.B, n, n, n), np.float32)
        p, dev = _ptr(out)
        self._ck(self.L.dsk_compute_grid_m(self.h, step, p, dev))
        return out

    def compute_grid_m_grad(self, step, g):
        p, dev = _ptr(_f32(g) if isinstance(g, np.ndarray) else g)
        self._ck(self.L.dsk_compute_grid_m_grad(self.h, step, p, dev))

    # ---- introspection -------------------------------------------------------------------------------
    def debug_cell_index(self, step, env=0):
        n = self.n_particles(env)
        base, key = np.zeros((n, 3), np.int32), np.zeros(n, np.int32)
        self._ck(self.L.dsk_debug_cell_index(self.h, step, env, C.c_void_p(base.ctypes.data), C.c_void_p(key.ctypes.data)))
        return base, key

    def debug_sort_order(self, env=0):
        perm = np.zeros(self.n_particles(env), np.int32)
        self._ck(self.L.dsk_debug_sort_order(self.h, env, C.c_void_p(perm.ctypes.data)))
        return perm

    def debug_grid(self, env=0, v_in=False, v_out=True, m=True, occupied=False):
        n = self.n_grid
        a = np.zeros((n, n, n, 3), np.float32) if v_in else None
        b = np.zeros((n, n, n, 3), np.float32) if v_out else None
        mm = np.zeros((n, n, n), np.float32) if m else None
        occ = np.zeros((n, n, n), np.uint8) if occupied else None
        self._ck(self.L.dsk_debug_grid(self.h, env, _ptr(a)[0], _ptr(b)[0], _ptr(mm)[0],
                                       C.c_void_p(occ.ctypes.data) if occupied else None))
        return a, b, mm, occ

    def debug_grid_grad(self, env=0):
        n = self.n_grid
        a, mm = np.zeros((n, n, n, 3), np.float32), np.zeros((n, n, n), np.float32)
        self._ck(self.L.dsk_debug_grid_grad(self.h, env, _ptr(a)[0], None, _ptr(mm)[0]))
        return a, mm

    def debug_frame(self, f, env=0):
        n = self.n_particles(env)
        x, v = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        F, Cm = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3, 3), np.float32)
        self._ck(self.L.dsk_debug_frame(self.h, f, env, _ptr(x)[0], _ptr(v)[0], _ptr(F)[0], _ptr(Cm)[0]))
        return x, v, F, Cm

    def debug_tool_frame(self, f, env, tool):
        s = np.zeros(8, np.float32)
        ci = np.full(max(1, len(self.scene.pairs)), -1, np.int32)
        self._ck(self.L.dsk_debug_tool_frame(self.h, f, env, tool, _ptr(s)[0], C.c_void_p(ci.ctypes.data)))
        return s, ci[:len(self.scene.pairs)]

    def debug_tool_frame_grad(self, f, env, tool):
        g = np.zeros(8, np.float32)
        self._ck(self.L.dsk_debug_tool_frame_grad(self.h, f, env, tool, _ptr(g)[0]))
        return g

    def debug_svd(self, F):
        F = _f32(F).reshape(-1, 9)
        n = len(F)
        U, s, V = np.zeros((n, 9), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 9), np.float32)
        self._ck(self.L.dsk_debug_svd(self.h, n, _ptr(F)[0], _ptr(U)[0], _ptr(s)[0], _ptr(V)[0]))
        return U.reshape(n, 3, 3), s, V.reshape(n, 3, 3)

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.dsk_launch_count(self.h, C.byref(n)))
        return n.value

    def memory_bytes(self):
        n = C.c_int64()
        self._ck(self.L.dsk_memory_bytes(self.h, C.byref(n)))
        return n.value

    # ---- measurement helpers -------------------------------------------------------------------------
    def set_graphs(self, on):
        self._ck(self.L.dsk_set_graphs(self.h, int(on)))

    def profile_enable(self, on):
        self._ck(self.L.dsk_profile_enable(self.h, int(on)))

    def kernel_classes(self):
        return [self.L.dsk_kernel_class_name(i).decode() for i in range(self.L.dsk_kernel_class_count())]

    def profile_report(self, reset=True):
        n = self.L.dsk_kernel_class_count()
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._ck(self.L.dsk_profile_report(self.h, ms, cnt, n, int(reset)))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.kernel_classes()) if cnt[i]}

    def timeline_enable(self, on=True):
        self._ck(self.L.dsk_timeline_enable(self.h, int(on)))

    def timeline_reset(self):
        self._ck(self.L.dsk_timeline_reset(self.h))

    def timeline_read(self, cap=16384):
        """[(kernel class, t0_ns, t1_ns)] of the launches stamped since the last reset, in launch order."""
        kid = (C.c_int * cap)()
        t0 = (C.c_ulonglong * cap)()
        t1 = (C.c_ulonglong * cap)()
        n = self.L.dsk_timeline_read(self.h, kid, t0, t1, cap)
        if n < 0:
            self._ck(n)
        names = self.kernel_classes()
        return [(names[kid[i]], int(t0[i]), int(t1[i])) for i in range(n) if t1[i] != 0]

    def launch_counts(self):
        n = self.L.dsk_kernel_class_count()
        cnt = (C.c_int64 * n)()
        self._ck(self.L.dsk_launch_counts(self.h, cnt, n))
        return {name: cnt[i] for i, name in enumerate(self.kernel_classes())}

    def loss_reset(self):
        self._ck(self.L.dsk_loss_reset(self.h))

    def loss_add_l2(self, step, target, weight=1.0):
        p, dev = _ptr(_f32(target) if isinstance(target, np.ndarray) else target)
        self._keep_t = target
        self._ck(self.L.dsk_loss_add_l2(self.h, step, p, C.c_double(weight), dev))

    # ---- multi-step fast path (one host call per phase of a rollout) -------------------------------------------
    def set_actions(self, step0, actions):
        """actions: [nsteps, B, A] numpy (host) or torch CUDA tensor."""
        a = _f32(actions) if isinstance(actions, np.ndarray) else actions
        n = int(a.shape[0])
        assert tuple(a.shape[1:]) == (self.B, self.A), f"actions must be [nsteps, {self.B}, {self.A}]"
        p, dev = _ptr(a)
        self._keep_a = a
        self._ck(self.L.dsk_set_actions(self.h, step0, n, p, dev))

    def forward_steps(self, step0, nsteps):
        self._ck(self.L.dsk_forward_steps(self.h, step0, nsteps))

    def backward_steps(self, step_hi, nsteps):
        self._ck(self.L.dsk_backward_steps(self.h, step_hi, nsteps))

    def loss_add_l2_steps(self, step0, nsteps, target, weight=1.0):
        p, dev = _ptr(_f32(target) if isinstance(target, np.ndarray) else target)
        self._keep_t = target
        self._ck(self.L.dsk_loss_add_l2_steps(self.h, step0, nsteps, p, C.c_double(weight), dev))

    def loss_get(self, out=None):
        if out is None:
            out = np.zeros(self.B, np.float32)
        p, dev = _ptr(out)
        self._ck(self.L.dsk_loss_get(self.h, p, dev))
        return out
