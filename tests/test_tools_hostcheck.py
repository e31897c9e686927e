"""CPU checks of the DEVICE tool math without a GPU.

`tests/host_check/host_check.cpp` compiles diffskill_b200/csrc/tools.cuh (the signed distance fields, normals, contact
response, forward kinematics and every hand-derived adjoint the grid / kinematics kernels call) with g++ and exposes it
through a small C ABI.  Here it is compared, tool by tool of every registered scene, with the oracle's restatement of
the reference code (plb/engine/primitive/primive_base.py:75-156, primitives.py) and its tape-AD adjoints.  Also the
reference's own property test for the tools (plb/engine/primitive/test_primitives.py:9-49: the analytic normal equals
the normalised gradient of the SDF) is run on the oracle.

The kernels around this math (launch geometry, shared memory, warp reductions, CUDA graphs) are covered by the
`-m gpu` parity tests only.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import ENVS, small_dough, tool_start
from diffskill_b200.engine import make_config
from diffskill_b200.scene import HAS_GAP, TOOL_BOX, TOOL_GRIPPER, TOOL_KNIFE
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
HC_DIR = os.path.join(HERE, 'host_check')
CSRC = os.path.join(os.path.dirname(HERE), 'diffskill_b200', 'csrc')
FP = C.POINTER(C.c_float)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(FP)


@pytest.fixture(scope='module')
def hc():
    so = os.path.join(HC_DIR, 'libhost_check.so')
    srcs = [os.path.join(HC_DIR, 'host_check.cpp'), os.path.join(HC_DIR, 'cuda_shim.h'),
            os.path.join(CSRC, 'tools.cuh'), os.path.join(CSRC, 'mpm_math.cuh')]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-w', '-o', so,
                               srcs[0]])
    L = C.CDLL(so)
    L.hc_tool_sdf.restype = C.c_float
    return L


def _setup(name):
    """fp32 + fp64 oracles of the scene with two consecutive, generic tool poses at frames 0 and 1."""
    scene, cfg, _ = small_dough(name, 8)
    cfgc = make_config(scene, 1, 8, 1, 1, True, 666., 0)
    rng = np.random.RandomState(7)
    st0 = [np.array(s, dtype=np.float64) for s in tool_start(name, scene)]
    st1 = []
    for t, s in zip(scene.tools, st0):
        q = s[3:7] + rng.normal(size=4) * 0.05          # generic, not unit norm
        if t.shape == 'Sphere':
            # the reference evaluates the Sphere's sdf / normal in world space (primitives.py:28-34) and ignores its
            # rotation there; the CUDA path works in the tool frame, which is the same for a unit quaternion only
            # (a Sphere's rotation is (1,0,0,0) unless a caller sets one)
            q = q / np.linalg.norm(q)
        s[3:7] = q
        n = s.copy()
        n[:3] += rng.normal(size=3) * 2e-4              # one substep of tool motion
        n[3:7] = q + rng.normal(size=4) * 2e-3
        n[3:7] /= np.linalg.norm(n[3:7])
        if t.type_id in HAS_GAP:
            n[7] = s[7] - 3e-4
        st1.append(n)
    st0 = [np.asarray(s, np.float32).astype(np.float64) for s in st0]
    st1 = [np.asarray(s, np.float32).astype(np.float64) for s in st1]
    oracles = []
    for f64 in (False, True):
        o = orc.Oracle(scene, 8, 3, f64=f64, threads=1)
        for i in range(len(scene.tools)):
            o.set_tool_state(0, i, st0[i])
            o.set_tool_state(1, i, st1[i])
        oracles.append(o)
    return scene, cfgc, st0, st1, oracles[0], oracles[1]


def _contact_points(o64, i, st, rng, want=40):
    """Points inside the soft-contact shell of tool i (influence > 0.1 <=> dist < ln(10)/666 = 3.5 mm) and a few inside."""
    out = []
    tries = 0
    while len(out) < want and tries < 200000:
        tries += 1
        p = st[:3] + rng.uniform(-1, 1, 3) * 0.3
        d = o64.tool_sdf(i, 0, p)
        if -0.004 < d < 0.003:
            out.append(np.asarray(p, np.float32).astype(np.float64))
    assert len(out) == want, f'could not sample the contact shell of tool {i}'
    return out


@pytest.mark.parametrize('name', ENVS)
def test_device_tool_math_matches_oracle(name, hc):
    scene, cfgc, st0, st1, o32, o64 = _setup(name)
    rng = np.random.RandomState(11)
    dt = C.c_float(scene.dt)
    for i, t in enumerate(scene.tools):
        desc = C.byref(cfgc.tools[i])
        fd_normals = t.type_id in (TOOL_BOX, TOOL_GRIPPER, TOOL_KNIFE)     # finite-difference normals: 1 ulp -> 3e-4
        keep = (_f(st0[i]), _f(st1[i]))
        p0, p1 = keep[0][1], keep[1][1]
        n_strict = 0
        pts = _contact_points(o64, i, st0[i], rng)
        for p in pts:
            v = rng.normal(size=3) * 0.5
            v = np.asarray(v, np.float32).astype(np.float64)
            pa, pp = _f(p)
            va, vp = _f(v)
            # ---- values --------------------------------------------------------------------------------------
            d_hc = hc.hc_tool_sdf(desc, p0, pp)
            assert abs(d_hc - o32.tool_sdf(i, 0, p)) <= 2e-7, (name, i, 'sdf')
            nrm = np.zeros(3, np.float32)
            hc.hc_tool_normal(desc, p0, pp, nrm.ctypes.data_as(FP))
            n32, n64 = o32.tool_normal(i, 0, p), o64.tool_normal(i, 0, p)
            assert np.abs(nrm - n64).max() <= max(3 * np.abs(n32 - n64).max(), 2e-6), (name, i, 'normal')
            out = np.zeros(3, np.float32)
            hc.hc_tool_collide(desc, p0, p1, pp, vp, dt, out.ctypes.data_as(FP))
            c32, c64 = o32.tool_collide(i, 0, p, v), o64.tool_collide(i, 0, p, v)
            sc = max(np.abs(c64).max(), 1e-3)
            assert np.abs(out - c64).max() <= max(3 * np.abs(c32 - c64).max(), 1e-5 * sc), (name, i, 'collide')
            # ---- adjoints: hand-derived device code against tape AD of the reference restatement ---------------
            for what, gdim in (('sdf', 1), ('normal', 3), ('collide', 3)):
                g = np.asarray(rng.normal(size=gdim), np.float32).astype(np.float64)
                g3 = np.zeros(3, np.float32)
                g3[:gdim] = g
                res = np.zeros(22, np.float32)
                hc.hc_tool_probe_grad(desc, p0, p1, {'sdf': 0, 'normal': 1, 'collide': 2}[what], pp, vp, dt,
                                      g3.ctypes.data_as(FP), res.ctypes.data_as(FP))
                r32 = np.concatenate(o32.tool_probe_grad(i, 0, what, p, v, g))
                r64 = np.concatenate(o64.tool_probe_grad(i, 0, what, p, v, g))
                if what != 'sdf':
                    # the hand adjoints do not return g(p) for the normal / contact (the grid node is not a variable)
                    res[0:3] = 0
                    r32[0:3] = 0
                    r64[0:3] = 0
                if t.shape == 'Sphere':
                    # tool-frame evaluation (see _setup): the adjoint of rotation[f] picks up a component along the
                    # quaternion itself (it changes |q|, which the reference's world-space normal never sees).  The
                    # normalising qmul of forward_kinematics annihilates exactly that direction one substep upstream,
                    # so action gradients are not affected; compare the tangential part.
                    qh = st0[i][3:7] / np.linalg.norm(st0[i][3:7])
                    for r_ in (res, r32, r64):
                        r_[9:13] -= np.dot(r_[9:13], qh) * qh
                for sl in (slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 13), slice(13, 14),
                           slice(14, 17), slice(17, 21), slice(21, 22)):
                    scale = max(np.abs(r64[sl]).max(), 1e-3 * max(np.abs(r64).max(), 1e-30))
                    e_hc = np.abs(res[sl] - r64[sl]).max()
                    e_32 = np.abs(r32[sl] - r64[sl]).max()
                    # collider velocity = pose difference / dt: pose adjoints of a contact are O(|g| / dt) and cancel
                    # to zero for separating nodes, so they carry an absolute rounding floor of a few ulp of |g| / dt
                    floor = 3e-7 * np.abs(g).max() / scene.dt if (what == 'collide' and sl.start >= 6) else 1e-12
                    assert e_hc <= max(3 * e_32, 2e-4 * scale) + floor, (name, i, t.shape, what, sl, res[sl], r64[sl])
                    n_strict += e_hc <= 2e-4 * scale + floor
        # the noise-floor escape hatch must stay the exception
        frac = n_strict / (len(pts) * 3 * 8)
        assert frac > (0.8 if fd_normals else 0.97), (name, i, t.shape, frac)


@pytest.mark.parametrize('name', ENVS)
def test_device_forward_kinematics_matches_oracle(name, hc):
    scene, cfgc, st0, st1, o32, o64 = _setup(name)
    rng = np.random.RandomState(13)
    for i, t in enumerate(scene.tools):
        desc = C.byref(cfgc.tools[i])
        for trial in range(20):
            a = np.asarray(rng.uniform(-1, 1, max(t.action_dim, 1)), np.float32)
            vel = np.zeros(7, np.float32)
            hc.hc_action_to_vel(desc, a.ctypes.data_as(FP), scene.substeps, vel.ctypes.data_as(FP))
            ref = np.zeros(7)
            if t.action_dim > 0:        # set_velocity, primive_base.py:260-268 (+ gap_vel, primitives.py:462-469)
                sc = np.asarray(t.action_scale, np.float32)
                ref[:3] = a[:3] * sc[:3] / np.float32(scene.substeps)
                if t.action_dim > 3:
                    ref[3:6] = a[3:6] * sc[3:6] / np.float32(scene.substeps)
                if t.type_id in HAS_GAP:
                    ref[6] = a[6] * sc[6] / np.float32(scene.substeps)
            assert np.abs(vel - ref).max() <= 1e-7 * max(np.abs(ref).max(), 1e-3)
            if trial % 2:               # also generic velocities in every slot (w of a 3-D tool is zero in practice)
                vel = np.asarray(rng.normal(size=7) * 2e-3, np.float32)
                if t.type_id not in HAS_GAP:
                    vel[6] = 0
            st = st0[i].copy()
            if trial >= 10:             # start on a position limit: the clamp's adjoint rule
                st[:3] = np.asarray(t.upper_bound if trial % 4 < 2 else t.lower_bound)
                st = np.asarray(st, np.float32).astype(np.float64)
            g = np.asarray(rng.normal(size=8), np.float32)
            if t.type_id not in HAS_GAP:
                g[7] = 0
            nxt = np.zeros(16, np.float32)
            gout = np.zeros(15, np.float32)
            sa, sp = _f(st)
            hc.hc_tool_fk(desc, sp, vel.ctypes.data_as(FP), nxt.ctypes.data_as(FP), g.ctypes.data_as(FP),
                          gout.ctypes.data_as(FP))
            n64, gs64, gv64 = o64.tool_probe_fk(i, st, vel, g)
            n32, gs32, gv32 = o32.tool_probe_fk(i, st, vel, g)
            assert np.abs(nxt[:8] - n64).max() <= max(3 * np.abs(n32 - n64).max(), 3e-7), (name, i, 'fk')
            assert np.abs(nxt[8:] - n64).max() <= max(3 * np.abs(n32 - n64).max(), 3e-7), (name, i, 'fk_inc')
            r, r32, r64 = gout, np.concatenate([gs32, gv32]), np.concatenate([gs64, gv64])
            if t.type_id not in HAS_GAP:
                # tools without a gap pass g(gap) straight through (the state slot is carried, not used)
                r[7] = r32[7] = r64[7] = 0
            for sl in (slice(0, 3), slice(3, 7), slice(7, 8), slice(8, 11), slice(11, 14), slice(14, 15)):
                scale = max(np.abs(r64[sl]).max(), 1e-6 * np.abs(r64).max())
                assert np.abs(r[sl] - r64[sl]).max() <= max(3 * np.abs(r32[sl] - r64[sl]).max(), 2e-4 * scale) + 1e-12, \
                    (name, i, t.shape, 'fk_adj', sl, r[sl], r64[sl])


ANALYTIC = {'Sphere', 'Capsule', 'RollingPin', 'RollingPinExt', 'Cylinder', 'Torus', 'Gripper2'}


@pytest.mark.parametrize('name', ENVS)
def test_analytic_normal_is_the_gradient_of_the_sdf(name):
    """The reference's own property test for the tools (plb/engine/primitive/test_primitives.py:9-49, run there on
    Cylinder with a random pose in f64, threshold 1e-5): `normal` == normalised gradient of `sdf`.  Here the gradient
    is the oracle's exact tape-AD gradient instead of the reference's central difference."""
    scene, cfgc, st0, st1, o32, o64 = _setup(name)
    rng = np.random.RandomState(17)
    for i, t in enumerate(scene.tools):
        if t.shape not in ANALYTIC:
            continue
        checked = 0
        st = st0[i].copy()
        st[3:7] /= np.linalg.norm(st[3:7])    # `normal` rotates by the raw quaternion: the property needs a unit one
        o64.set_tool_state(1, i, st)
        for _ in range(300):
            p = st[:3] + rng.uniform(-1, 1, 3) * 0.25
            g = o64.tool_probe_grad(i, 1, 'sdf', p, np.zeros(3), [1.0])[0]
            n = o64.tool_normal(i, 1, p)
            gn = g / np.linalg.norm(g)
            nn = n / np.linalg.norm(n)
            # skip the measure-zero kinks (cylinder edge / axis, torus centre circle) where the two branch differently
            eps = 1e-6
            g2 = o64.tool_probe_grad(i, 1, 'sdf', p + eps * gn, np.zeros(3), [1.0])[0]
            if np.abs(g2 - g).max() > 1e-2:
                continue
            assert np.abs(gn - nn).max() < 1e-5, (name, t.shape, p, gn, nn)
            checked += 1
        assert checked > 250
