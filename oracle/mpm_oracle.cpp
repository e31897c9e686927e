// oracle/mpm_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see mpm_oracle.h header).
//
// Literal CPU restatement of the reference's differentiable MLS-MPM hot path.
// One C++ function per reference Taichi kernel, same fields, same schedule:
//   substep       plb/engine/mpm_simulator.py:307-323
//   substep_grad  plb/engine/mpm_simulator.py:325-345
// The forward arithmetic is written once, templated on the scalar type S, and
// instantiated with S = T (plain forward) and S = ad::Var<T> (tape recording);
// every `.grad` kernel replays the forward of ONE particle / node / tool on the
// tape and reverses it with Taichi's per-op adjoint rules (oracle/ad.hpp).
// `svd_grad` is the reference's hand-written formula (mpm_simulator.py:131-156).
//
// T = float reproduces the reference's fp32 run (SIMULATOR.dtype float32,
// plb/config/default_config.py:15); T = double is the twin used to measure the
// fp32 noise floor and to finite-difference the adjoints.
//
// PARITY UNPINNED (no taichi offline, no golden vectors in the reference).
// ti.svd is restated as a one-sided Jacobi SVD with U,V proper rotations,
// singular values sorted by decreasing magnitude, sign on the last one.
//
// Compile with -ffp-contract=off so every +,-,*,/ rounds separately.
#include "mpm_oracle.h"

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ad.hpp"

namespace orc {
using namespace ad;

// ----------------------------------------------------------------------------
// small fixed-size algebra (Taichi evaluation order: left-to-right sums)
// ----------------------------------------------------------------------------
template <class S>
struct V3 {
  S a[3];
  S& operator[](int i) { return a[i]; }
  const S& operator[](int i) const { return a[i]; }
};
template <class S>
struct Q4 {
  S a[4];
  S& operator[](int i) { return a[i]; }
  const S& operator[](int i) const { return a[i]; }
};
template <class S>
struct M3 {
  S a[3][3];
};

template <class S>
inline V3<S> operator+(const V3<S>& x, const V3<S>& y) {
  return V3<S>{{x[0] + y[0], x[1] + y[1], x[2] + y[2]}};
}
template <class S>
inline V3<S> operator-(const V3<S>& x, const V3<S>& y) {
  return V3<S>{{x[0] - y[0], x[1] - y[1], x[2] - y[2]}};
}
template <class S, class R>
inline V3<S> scale(const V3<S>& x, const R& s) {
  return V3<S>{{x[0] * s, x[1] * s, x[2] * s}};
}
template <class S>
inline S dot(const V3<S>& x, const V3<S>& y) {
  return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}
template <class S>
inline V3<S> cross(const V3<S>& a, const V3<S>& b) {
  return V3<S>{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}
template <class S>
inline M3<S> matmul(const M3<S>& A, const M3<S>& B) {
  M3<S> C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      S acc = A.a[i][0] * B.a[0][j];
      acc = acc + A.a[i][1] * B.a[1][j];
      acc = acc + A.a[i][2] * B.a[2][j];
      C.a[i][j] = acc;
    }
  return C;
}
template <class S>
inline M3<S> transpose(const M3<S>& A) {
  M3<S> C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.a[i][j] = A.a[j][i];
  return C;
}
template <class S>
inline V3<S> matvec(const M3<S>& A, const V3<S>& x) {
  V3<S> y;
  for (int i = 0; i < 3; i++) {
    S acc = A.a[i][0] * x[0];
    acc = acc + A.a[i][1] * x[1];
    acc = acc + A.a[i][2] * x[2];
    y[i] = acc;
  }
  return y;
}
template <class S>
inline S det3(const M3<S>& m) {
  const auto& a = m.a;
  return a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[1][0] * (a[0][1] * a[2][2] - a[2][1] * a[0][2]) +
         a[2][0] * (a[0][1] * a[1][2] - a[1][1] * a[0][2]);
}

// ----------------------------------------------------------------------------
// quaternion helpers  (plb/engine/primitive/utils.py:4-54), (w,x,y,z)
// ----------------------------------------------------------------------------
template <class S>
inline V3<S> qrot(const Q4<S>& rot, const V3<S>& v) {  // utils.py:9-15
  typedef typename real_of<S>::type T;
  V3<S> qv{{rot[1], rot[2], rot[3]}};
  V3<S> uv = cross(qv, v);
  V3<S> uuv = cross(qv, uv);
  V3<S> r;
  for (int i = 0; i < 3; i++) r[i] = v[i] + T(2) * (rot[0] * uv[i] + uuv[i]);
  return r;
}
template <class S>
inline Q4<S> qnormalized(const Q4<S>& q) {  // ti Vector.normalized(): 1/(norm) * q
  typedef typename real_of<S>::type T;
  S n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  S inv = T(1) / s_sqrt(n2);
  return Q4<S>{{inv * q[0], inv * q[1], inv * q[2], inv * q[3]}};
}
template <class S>
inline Q4<S> qconj_normalized(const Q4<S>& q) {
  Q4<S> c{{q[0], -q[1], -q[2], -q[3]}};
  return qnormalized(c);
}
template <class S>
inline Q4<S> qmul(const Q4<S>& q, const Q4<S>& r) {  // utils.py:23-31 (normalising)
  typedef typename real_of<S>::type T;
  // terms[i][j] = r[i]*q[j]
  S w = r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3];
  S x = r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2];
  S y = r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1];
  S z = r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0];
  S n = s_sqrt(w * w + x * x + y * y + z * z);
  (void)sizeof(T);
  return Q4<S>{{w / n, x / n, y / n, z / n}};
}
template <class S>
inline Q4<S> w2quat(const V3<S>& aa) {  // utils.py:34-47
  typedef typename real_of<S>::type T;
  S w = s_sqrt(dot(aa, aa) + T(1e-16));
  Q4<S> out{{S(T(1)), S(T(0)), S(T(0)), S(T(0))}};
  if (val(w) > T(1e-9)) {
    S sn = s_sin(w / T(2));
    out[0] = s_cos(w / T(2));
    out[1] = (aa[0] / w) * sn;
    out[2] = (aa[1] / w) * sn;
    out[3] = (aa[2] / w) * sn;
  }
  return out;
}
template <class S>
inline V3<S> inv_trans(const V3<S>& p, const V3<S>& position, const Q4<S>& rotation) {  // utils.py:50-54
  return qrot(qconj_normalized(rotation), p - position);
}
template <class S>
inline S length8(const V3<S>& x) {  // utils.py:4-6
  typedef typename real_of<S>::type T;
  return s_sqrt(dot(x, x) + T(1e-8));
}
template <class S>
inline S length14(const V3<S>& x) {  // primitives.py:13-15
  typedef typename real_of<S>::type T;
  return s_sqrt(dot(x, x) + T(1e-14));
}

// ----------------------------------------------------------------------------
// tools
// ----------------------------------------------------------------------------
inline bool is_gripper(int type) {  // tools with two jaws applied one after the other (primitives.py:428, :576)
  return type == ORC_TOOL_GRIPPER || type == ORC_TOOL_GRIPPER2;
}
inline bool has_gap(int type) {  // tools with a gap state, an 8-float state and a 7-D action (+ Chopsticks, primitives.py:218)
  return is_gripper(type) || type == ORC_TOOL_CHOPSTICKS;
}
template <class T>
struct ToolC {  // constants rounded to T as Taichi rounds python floats into fields/constants
  int type, action_dim;
  T action_scale[8];
  T friction, softness;
  T lo[3], hi[3];
  T size[3];
  T h, half_h, r, radius;
  T prism_h[2], prot[4];
  T min_gap, max_gap;
};

template <class S>
struct Pose {
  V3<S> pos;
  Q4<S> rot;
  S gap;
};

template <class S, class T>
inline S box_sdf(const T size[3], const V3<S>& p) {  // primitives.py:374-380
  V3<S> q{{s_abs(p[0]) - size[0], s_abs(p[1]) - size[1], s_abs(p[2]) - size[2]}};
  V3<S> mq{{s_max(q[0], T(0)), s_max(q[1], T(0)), s_max(q[2], T(0))}};
  S out = length14(mq);
  out = out + s_min(s_max(q[0], s_max(q[1], q[2])), T(0));
  return out;
}
template <class S, class T>
inline V3<S> capsule_p2(const ToolC<T>& c, const V3<S>& p) {  // primitives.py:56-58
  V3<S> p2 = p;
  p2[1] = p2[1] + c.half_h;
  p2[1] = p2[1] - s_min(s_max(p2[1], T(0)), c.h);
  return p2;
}
template <class S, class T>
inline S prism_sdf(const ToolC<T>& c, const V3<S>& p0) {  // primitives.py:711-718
  Q4<S> pr{{S(c.prot[0]), S(c.prot[1]), S(c.prot[2]), S(c.prot[3])}};
  V3<S> p = qrot(qconj_normalized(pr), p0);
  S q0 = s_abs(p[0]), q2 = s_abs(p[2]);
  return s_max(q2 - c.prism_h[1], s_max(q0 * T(0.866025) + p[1] * T(0.5), -p[1]) - c.prism_h[0] * T(0.5));
}
template <class S>
inline S length14_2(const S& a, const S& b) {  // primitives.py:13-15 on a 2-vector
  typedef typename real_of<S>::type T;
  return s_sqrt(a * a + b * b + T(1e-14));
}
template <class S, class T>
inline S cylinder_sdf(const ToolC<T>& c, const V3<S>& p) {  // primitives.py:309-313 (h = radial, r = axial half extent)
  S l = length14_2(p[0], p[2]);
  S d0 = s_abs(l) - c.h, d1 = s_abs(p[1]) - c.r;
  return s_min(s_max(d0, d1), T(0)) + length14_2(s_max(d0, T(0)), s_max(d1, T(0)));
}
template <class S, class T>
inline V3<S> cylinder_normal(const ToolC<T>& c, const V3<S>& p) {  // primitives.py:315-329
  S l = length14_2(p[0], p[2]);
  S d0 = l - c.h, d1 = s_abs(p[1]) - c.r;
  T f = val(d0) > val(d1) ? T(1) : T(0);                       // casts: no gradient
  T inside = val(s_max(d0, d1)) <= T(0) ? T(1) : T(0);
  S n20 = s_max(d0, T(0)) + inside * f, n21 = s_max(d1, T(0)) + inside * (T(1) - f);
  S ln = length14_2(n20, n21);
  S a = n20 / ln, b = n21 / ln;
  S p20 = p[0] / l, p21 = p[2] / l;
  T sy = (val(p[1]) >= T(0) ? T(1) : T(0)) * T(2) - T(1);
  V3<S> n3{{p20 * a, b * sy, p21 * a}};
  S l3 = length14(n3);
  return V3<S>{{n3[0] / l3, n3[1] / l3, n3[2] / l3}};
}
template <class S, class T>
inline S torus_sdf(const ToolC<T>& c, const V3<S>& p) {  // primitives.py:344-347 (tx in h, ty in r)
  S q0 = length14_2(p[0], p[2]) - c.h;
  return length14_2(q0, p[1]) - c.r;
}
template <class S, class T>
inline V3<S> torus_normal(const ToolC<T>& c, const V3<S>& p) {  // primitives.py:349-358
  S l = length14_2(p[0], p[2]);
  S q0 = length14_2(p[0], p[2]) - c.h, q1 = p[1];
  S lq = length14_2(q0, q1);
  S n20 = q0 / lq, n21 = q1 / lq;
  S x20 = p[0] / l, x21 = p[2] / l;
  V3<S> n3{{x20 * n20, n21, x21 * n20}};
  S l3 = length14(n3);
  return V3<S>{{n3[0] / l3, n3[1] / l3, n3[2] / l3}};
}
template <class S, class T>
inline S local_sdf(const ToolC<T>& c, const V3<S>& p) {
  switch (c.type) {
    case ORC_TOOL_CAPSULE:
    case ORC_TOOL_ROLLINGPIN_EXT:
    case ORC_TOOL_ROLLINGPIN:
    case ORC_TOOL_GRIPPER2:
      return length14(capsule_p2(c, p)) - c.r;  // primitives.py:59
    case ORC_TOOL_CYLINDER:
      return cylinder_sdf(c, p);
    case ORC_TOOL_TORUS:
      return torus_sdf(c, p);
    case ORC_TOOL_BOX:
    case ORC_TOOL_GRIPPER:
      return box_sdf<S, T>(c.size, p);
    case ORC_TOOL_KNIFE:
      return s_max(prism_sdf(c, p), box_sdf<S, T>(c.size, p));  // primitives.py:769-773
    default:
      return S(T(0));
  }
}
template <class S, class T>
inline V3<S> local_normal(const ToolC<T>& c, const V3<S>& p) {
  if (c.type == ORC_TOOL_CYLINDER) return cylinder_normal(c, p);
  if (c.type == ORC_TOOL_TORUS) return torus_normal(c, p);
  if (c.type == ORC_TOOL_CAPSULE || c.type == ORC_TOOL_ROLLINGPIN_EXT || c.type == ORC_TOOL_ROLLINGPIN ||
      c.type == ORC_TOOL_GRIPPER2) {  // primitives.py:61-66
    V3<S> p2 = capsule_p2(c, p);
    S l = length14(p2);
    return V3<S>{{p2[0] / l, p2[1] / l, p2[2] / l}};
  }
  // central finite differences, primitives.py:382-393 / 796-807
  T d = (T)(float)1e-4;
  V3<S> n;
  for (int i = 0; i < 3; i++) {
    V3<S> inc = p, dec = p;
    inc[i] = inc[i] + d;
    dec[i] = dec[i] - d;
    n[i] = (T(0.5) / d) * (local_sdf(c, inc) - local_sdf(c, dec));
  }
  S l = length14(n);
  return V3<S>{{n[0] / l, n[1] / l, n[2] / l}};
}
template <class S>
inline V3<S> gripper_pos(const Pose<S>& P, int flag) {  // primitives.py:471-473
  typedef typename real_of<S>::type T;
  V3<S> off{{P.gap / T(2) * T(flag), S(T(0)), S(T(0))}};
  return P.pos + qrot(P.rot, off);
}
template <class S, class T>
inline S sdf2(const ToolC<T>& c, const Pose<S>& P, const V3<S>& p, int flag) {  // primitives.py:475-478
  if (c.type == ORC_TOOL_GRIPPER2)  // primitives.py:607-610: Capsule._sdf
    return length14(capsule_p2(c, inv_trans(p, gripper_pos(P, flag), P.rot))) - c.r;
  return box_sdf<S, T>(c.size, inv_trans(p, gripper_pos(P, flag), P.rot));
}
template <class S, class T>
inline V3<S> normal2(const ToolC<T>& c, const Pose<S>& P, const V3<S>& p, int flag) {  // primitives.py:480-483
  ToolC<T> b = c;
  b.type = c.type == ORC_TOOL_GRIPPER2 ? ORC_TOOL_CAPSULE : ORC_TOOL_BOX;  // primitives.py:612-615: Capsule._normal
  return qrot(P.rot, local_normal(b, inv_trans(p, gripper_pos(P, flag), P.rot)));
}
// Chopsticks (primitives.py:245-261): two capsules at -+gap/2 along local x, shifted by h/2 along local y, inside the ONE
// tool frame; sdf = min of the two, normal = the nearer one's (a <= b)
template <class S, class T>
inline void chopsticks_points(const ToolC<T>& c, const Pose<S>& P, const V3<S>& p, V3<S>& pa, V3<S>& pb) {
  V3<S> g = inv_trans(p, P.pos, P.rot);
  V3<S> q{{g[0], g[1] - (-c.h / T(2)), g[2]}};   // grid_pos - (0, -h/2, 0)
  S half = P.gap / T(2);
  pa = V3<S>{{q[0] - half, q[1], q[2]}};          // p - delta
  pb = V3<S>{{q[0] + half, q[1], q[2]}};          // p + delta
}
template <class S, class T>
inline S tool_sdf(const ToolC<T>& c, const Pose<S>& P, const V3<S>& p) {
  if (c.type == ORC_TOOL_CHOPSTICKS) {
    V3<S> pa, pb;
    chopsticks_points(c, P, p, pa, pb);
    return s_min(length14(capsule_p2(c, pa)) - c.r, length14(capsule_p2(c, pb)) - c.r);
  }
  if (is_gripper(c.type)) return s_min(sdf2(c, P, p, -1), sdf2(c, P, p, 1));  // primitives.py:485-487
  if (c.type == ORC_TOOL_SPHERE) return length14(p - P.pos) - c.radius;               // primitives.py:28-30
  return local_sdf(c, inv_trans(p, P.pos, P.rot));                                      // primive_base.py:75-78
}
template <class S, class T>
inline V3<S> tool_normal(const ToolC<T>& c, const Pose<S>& P, const V3<S>& p) {
  if (c.type == ORC_TOOL_CHOPSTICKS) {  // primitives.py:253-261
    V3<S> pa, pb;
    chopsticks_points(c, P, p, pa, pb);
    V3<S> a2 = capsule_p2(c, pa), b2 = capsule_p2(c, pb);
    S la = length14(a2), lb = length14(b2);
    T m = val(la - c.r) <= val(lb - c.r) ? T(1) : T(0);
    V3<S> r;
    for (int i = 0; i < 3; i++) r[i] = m * (a2[i] / la) + (T(1) - m) * (b2[i] / lb);
    return qrot(P.rot, r);
  }
  if (is_gripper(c.type)) {  // primitives.py:489-496
    S a = sdf2(c, P, p, -1), b = sdf2(c, P, p, 1);
    V3<S> an = normal2(c, P, p, -1), bn = normal2(c, P, p, 1);
    T m = val(a) <= val(b) ? T(1) : T(0);
    V3<S> r;
    for (int i = 0; i < 3; i++) r[i] = m * an[i] + (T(1) - m) * bn[i];
    return r;
  }
  if (c.type == ORC_TOOL_SPHERE) {  // primitives.py:32-34
    V3<S> d = p - P.pos;
    S l = length14(d);
    return V3<S>{{d[0] / l, d[1] / l, d[2] / l}};
  }
  return qrot(P.rot, local_normal(c, inv_trans(p, P.pos, P.rot)));  // primive_base.py:80-85
}

// the friction/soft-contact response shared by Primitive.collide (eps 1e-8, primive_base.py:96-120)
// and Gripper.collide2 (eps 1e-14, primitives.py:513-536)
template <class S, class T>
inline V3<S> contact_response(const V3<S>& v_out, const V3<S>& D, const V3<S>& cv, const S& influence, T friction,
                              bool eps14) {
  V3<S> input_v = v_out - cv;
  S nc = dot(input_v, D);
  S mn = s_min(nc, T(0));
  V3<S> t = input_v - scale(D, mn);
  S tn = eps14 ? length14(t) : length8(t);
  S mx = s_max(T(0), tn + nc * friction);
  V3<S> tf{{t[0] / tn * mx, t[1] / tn * mx, t[2] / tn * mx}};
  T flag = (val(nc) < T(0) && std::sqrt(val(dot(t, t))) > T(1e-30)) ? T(1) : T(0);
  V3<S> r;
  for (int i = 0; i < 3; i++) {
    S t2 = tf[i] * flag + t[i] * (T(1) - flag);
    r[i] = cv[i] + input_v[i] * (T(1) - influence) + t2 * influence;
  }
  return r;
}

template <class S, class T>
inline V3<S> tool_collide(const ToolC<T>& c, const Pose<S>& P0, const Pose<S>& P1, const V3<S>& p, V3<S> v_out, T dt) {
  if (is_gripper(c.type)) {  // primitives.py:507-536
    for (int flag = -1; flag <= 1; flag += 2) {
      S dist = sdf2(c, P0, p, flag);
      S influence = s_min(s_exp(-dist * c.softness), T(1));
      if ((c.softness > T(0) && val(influence) > T(0.1)) || val(dist) <= T(0)) {
        V3<S> D = normal2(c, P0, p, flag);
        // collider_v, primitives.py:498-505
        V3<S> rel = qrot(qconj_normalized(P0.rot), p - gripper_pos(P0, flag));
        V3<S> np = qrot(P1.rot, rel) + gripper_pos(P1, flag);
        V3<S> cv{{(np[0] - p[0]) / dt, (np[1] - p[1]) / dt, (np[2] - p[2]) / dt}};
        v_out = contact_response<S, T>(v_out, D, cv, influence, c.friction, true);
      }
    }
    return v_out;
  }
  // primive_base.py:96-120
  S dist = tool_sdf(c, P0, p);
  S influence = s_min(s_exp(-dist * c.softness), T(1));
  if ((c.softness > T(0) && val(influence) > T(0.1)) || val(dist) <= T(0)) {
    V3<S> D = tool_normal(c, P0, p);
    // collider_v, primive_base.py:87-94
    V3<S> rel = qrot(qconj_normalized(P0.rot), p - P0.pos);
    V3<S> np = qrot(P1.rot, rel) + P1.pos;
    V3<S> cv{{(np[0] - p[0]) / dt, (np[1] - p[1]) / dt, (np[2] - p[2]) / dt}};
    v_out = contact_response<S, T>(v_out, D, cv, influence, c.friction, false);
  }
  return v_out;
}

// forward kinematics of one tool for one substep
template <class S, class T>
inline Pose<S> tool_fk(const ToolC<T>& c, const Pose<S>& P, const V3<S>& v, const V3<S>& w, const S& gap_vel) {
  Pose<S> N;
  N.gap = P.gap;
  V3<S> step = v;
  if (c.type == ORC_TOOL_ROLLINGPIN_EXT || c.type == ORC_TOOL_ROLLINGPIN) {  // primitives.py:120-136 / :101-117
    S dw = v[0], dth = v[1], dy = v[2];
    V3<S> down{{S(T(0)), S(T(-1)), S(T(0))}};
    V3<S> y_dir = qrot(P.rot, down);
    V3<S> up{{S(T(0)), S(T(1)), S(T(0))}};
    V3<S> x_dir = c.type == ORC_TOOL_ROLLINGPIN ? scale(scale(cross(up, y_dir), dw), T(0.03))  // :110
                                                : scale(cross(up, y_dir), dw * T(0.03) + w[0]);
    x_dir[1] = dy;
    V3<S> a1{{S(T(0)), -dth, S(T(0))}}, a2{{S(T(0)), dw, S(T(0))}};
    N.rot = qmul(w2quat(a1), qmul(P.rot, w2quat(a2)));
    step = x_dir;
  } else if (c.type == ORC_TOOL_CHOPSTICKS) {  // primitives.py:230-234: no upper clamp on the gap
    N.gap = s_max(P.gap - gap_vel, c.min_gap);
    N.rot = qmul(P.rot, w2quat(w));
  } else if (is_gripper(c.type)) {  // primitives.py:456-460
    N.gap = s_min(s_max(P.gap - gap_vel, c.min_gap), c.max_gap);
    N.rot = qmul(P.rot, w2quat(w));
  } else {  // primive_base.py:152-156
    N.rot = qmul(w2quat(w), P.rot);
  }
  for (int k = 0; k < 3; k++) N.pos[k] = s_max(s_min(P.pos[k] + step[k], c.hi[k]), c.lo[k]);
  return N;
}

// ----------------------------------------------------------------------------
// 3x3 SVD restating ti.svd (third-party intrinsic, call site mpm_simulator.py:129)
// ----------------------------------------------------------------------------
template <class T>
void svd3(const T A[9], T U[9], T sig[3], T V[9]) {
  T B[3][3], W[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      B[i][j] = A[i * 3 + j];
      W[i][j] = i == j ? T(1) : T(0);
    }
  const int sweeps = sizeof(T) == 4 ? 4 : 9;
  const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  for (int sw = 0; sw < sweeps; sw++)
    for (int k = 0; k < 3; k++) {
      int p = PQ[k][0], q = PQ[k][1];
      T al = B[0][p] * B[0][p] + B[1][p] * B[1][p] + B[2][p] * B[2][p];
      T be = B[0][q] * B[0][q] + B[1][q] * B[1][q] + B[2][q] * B[2][q];
      T ga = B[0][p] * B[0][q] + B[1][p] * B[1][q] + B[2][p] * B[2][q];
      if (ga == T(0)) continue;
      T zeta = (be - al) / (T(2) * ga);
      T t = (zeta >= T(0) ? T(1) : T(-1)) / (std::fabs(zeta) + std::sqrt(T(1) + zeta * zeta));
      T c = T(1) / std::sqrt(T(1) + t * t), s = c * t;
      for (int i = 0; i < 3; i++) {
        T bp = B[i][p], bq = B[i][q];
        B[i][p] = c * bp - s * bq;
        B[i][q] = s * bp + c * bq;
        T vp = W[i][p], vq = W[i][q];
        W[i][p] = c * vp - s * vq;
        W[i][q] = s * vp + c * vq;
      }
    }
  T n2[3];
  for (int j = 0; j < 3; j++) n2[j] = B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j];
  auto swapneg = [&](int p, int q) {  // swap columns p,q and negate the new q: keeps det(V)=+1, B=A V
    for (int i = 0; i < 3; i++) {
      T b = B[i][p];
      B[i][p] = B[i][q];
      B[i][q] = -b;
      T w = W[i][p];
      W[i][p] = W[i][q];
      W[i][q] = -w;
    }
    std::swap(n2[p], n2[q]);
  };
  if (n2[0] < n2[1]) swapneg(0, 1);
  if (n2[0] < n2[2]) swapneg(0, 2);
  if (n2[1] < n2[2]) swapneg(1, 2);
  T s0 = std::sqrt(n2[0]), s1 = std::sqrt(n2[1]);
  T u0[3], u1[3], u2[3];
  if (s0 > T(0)) {
    for (int i = 0; i < 3; i++) u0[i] = B[i][0] / s0;
  } else {
    u0[0] = 1;
    u0[1] = 0;
    u0[2] = 0;
  }
  if (s1 > T(1e-18)) {
    for (int i = 0; i < 3; i++) u1[i] = B[i][1] / s1;
  } else {  // rank <= 1: any unit vector orthogonal to u0
    int k = std::fabs(u0[0]) < std::fabs(u0[1]) ? (std::fabs(u0[0]) < std::fabs(u0[2]) ? 0 : 2)
                                                  : (std::fabs(u0[1]) < std::fabs(u0[2]) ? 1 : 2);
    T e[3] = {0, 0, 0};
    e[k] = 1;
    T d = u0[k];
    T nrm = 0;
    for (int i = 0; i < 3; i++) {
      u1[i] = e[i] - d * u0[i];
      nrm += u1[i] * u1[i];
    }
    nrm = std::sqrt(nrm);
    for (int i = 0; i < 3; i++) u1[i] /= nrm;
  }
  u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
  u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
  u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
  sig[0] = s0;
  sig[1] = s1;
  sig[2] = u2[0] * B[0][2] + u2[1] * B[1][2] + u2[2] * B[2][2];
  for (int i = 0; i < 3; i++) {
    U[i * 3 + 0] = u0[i];
    U[i * 3 + 1] = u1[i];
    U[i * 3 + 2] = u2[i];
    for (int j = 0; j < 3; j++) V[i * 3 + j] = W[i][j];
  }
}

// Second, independent SVD algorithm (test infrastructure for the parity policy): the construction of the
// McAdams-Selle-Tamstorf-Teran-Sifakis 3x3 SVD that `ti.svd` is recalled to use -- cyclic Jacobi
// EIGEN-decomposition of the symmetric S = A^T A (exact Givens angles here, the original approximates them), columns of
// B = A V sorted by decreasing norm, then a Givens QR of B: U = Q, sigma = diag(R), sign carried by the last value.
// It shares no code with svd3 above: different iteration (two-sided on S instead of one-sided on B), different U
// (QR instead of normalised columns).  orc_set_svd_algorithm(1) makes every substep use it, so that forward states and
// -- through backward_svd's 1/clamp(sigma_j^2 - sigma_i^2) -- gradients can be compared ACROSS algorithms, in particular
// in degenerate singular subspaces (F = I at rest) where U and V are a choice.
int g_svd_algorithm = 0;
template <class T>
void svd3_eig_qr(const T A[9], T U[9], T sig[3], T V[9]) {
  T S[3][3], W[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      S[i][j] = A[0 * 3 + i] * A[0 * 3 + j] + A[1 * 3 + i] * A[1 * 3 + j] + A[2 * 3 + i] * A[2 * 3 + j];
      W[i][j] = i == j ? T(1) : T(0);
    }
  const int sweeps = sizeof(T) == 4 ? 5 : 10;
  const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  for (int sw = 0; sw < sweeps; sw++)
    for (int k = 0; k < 3; k++) {
      int p = PQ[k][0], q = PQ[k][1];
      T apq = S[p][q];
      if (apq == T(0)) continue;
      T theta = (S[q][q] - S[p][p]) / (T(2) * apq);
      T t = (theta >= T(0) ? T(1) : T(-1)) / (std::fabs(theta) + std::sqrt(T(1) + theta * theta));
      T c = T(1) / std::sqrt(T(1) + t * t), s = c * t;
      // S <- J^T S J, W <- W J with J = [[c, s], [-s, c]] on (p, q)
      for (int i = 0; i < 3; i++) {
        T sp = S[i][p], sq = S[i][q];
        S[i][p] = c * sp - s * sq;
        S[i][q] = s * sp + c * sq;
      }
      for (int j = 0; j < 3; j++) {
        T sp = S[p][j], sq = S[q][j];
        S[p][j] = c * sp - s * sq;
        S[q][j] = s * sp + c * sq;
      }
      for (int i = 0; i < 3; i++) {
        T wp = W[i][p], wq = W[i][q];
        W[i][p] = c * wp - s * wq;
        W[i][q] = s * wp + c * wq;
      }
    }
  T B[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[i][j] = A[i * 3 + 0] * W[0][j] + A[i * 3 + 1] * W[1][j] + A[i * 3 + 2] * W[2][j];
  T n2[3];
  for (int j = 0; j < 3; j++) n2[j] = B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j];
  auto swapneg = [&](int p, int q) {
    for (int i = 0; i < 3; i++) {
      T b = B[i][p];
      B[i][p] = B[i][q];
      B[i][q] = -b;
      T w = W[i][p];
      W[i][p] = W[i][q];
      W[i][q] = -w;
    }
    std::swap(n2[p], n2[q]);
  };
  if (n2[0] < n2[1]) swapneg(0, 1);
  if (n2[0] < n2[2]) swapneg(0, 2);
  if (n2[1] < n2[2]) swapneg(1, 2);
  // Givens QR of B: Q accumulates the rotations, R = Q^T B ends upper triangular (diagonal up to round-off)
  T Q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  auto givens = [&](int a, int b, int col) {   // zero B[b][col] against B[a][col]
    T x = B[a][col], y = B[b][col];
    T r = std::sqrt(x * x + y * y);
    if (r == T(0)) return;
    T c = x / r, s = y / r;
    for (int j = 0; j < 3; j++) {
      T ba = B[a][j], bb = B[b][j];
      B[a][j] = c * ba + s * bb;
      B[b][j] = -s * ba + c * bb;
    }
    for (int i = 0; i < 3; i++) {
      T qa = Q[i][a], qb = Q[i][b];
      Q[i][a] = c * qa + s * qb;
      Q[i][b] = -s * qa + c * qb;
    }
  };
  givens(0, 1, 0);
  givens(0, 2, 0);
  givens(1, 2, 1);
  // r00, r11 >= 0 by construction; the sign of det(A) ends up in r22
  sig[0] = B[0][0];
  sig[1] = B[1][1];
  sig[2] = B[2][2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      U[i * 3 + j] = Q[i][j];
      V[i * 3 + j] = W[i][j];
    }
}

// ----------------------------------------------------------------------------
// per-element forward bodies, templated on S
// ----------------------------------------------------------------------------
template <class T>
struct Consts {
  int n;
  T dt, dx, inv_dx, p_mass, c_stress, c_C, x_hi, x_lo, m_eps, ground_friction;
  double ground_friction_d;
  T gravity[3];
};

template <class S, class T>
inline void bspline(const Consts<T>& k, const V3<S>& x, int base[3], V3<S>& fx, V3<S> w[3]) {
  // mpm_simulator.py:201-204 : trunc cast, separate roundings
  for (int d = 0; d < 3; d++) {
    S xg = x[d] * k.inv_dx;
    base[d] = (int)(val(xg) - T(0.5));
    fx[d] = xg - T(base[d]);
    S a = T(1.5) - fx[d], b = fx[d] - T(1), c = fx[d] - T(0.5);
    w[0][d] = T(0.5) * (a * a);
    w[1][d] = T(0.75) - b * b;
    w[2][d] = T(0.5) * (c * c);
  }
}

// compute_von_mises, mpm_simulator.py:165-182 ; sig is the diagonal
template <class S, class T>
inline M3<S> von_mises(const M3<S>& F, const M3<S>& U, const V3<S>& sig_in, const M3<S>& V, T yield_stress, T mu) {
  V3<S> sig{{s_max(sig_in[0], T(0.05)), s_max(sig_in[1], T(0.05)), s_max(sig_in[2], T(0.05))}};
  V3<S> eps{{s_log(sig[0]), s_log(sig[1]), s_log(sig[2])}};
  S mean = (eps[0] + eps[1] + eps[2]) / T(3);
  V3<S> eh{{eps[0] - mean, eps[1] - mean, eps[2] - mean}};
  S ehn = s_sqrt(dot(eh, eh) + T(1e-8));
  S dg = ehn - yield_stress / (T(2) * mu);
  if (val(dg) > T(0)) {
    S kf = dg / ehn;
    M3<S> US, R;
    S e[3];
    for (int i = 0; i < 3; i++) e[i] = s_exp(eps[i] - kf * eh[i]);
    // U @ diag(e) @ V^T with Taichi's matmul order (zeros of the diagonal matrix add exactly)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) US.a[i][j] = U.a[i][j] * e[j];
    R = matmul(US, transpose(V));
    return R;
  }
  return F;
}

template <class S, class T>
struct P2GOut {
  M3<S> newF;
  V3<S> mv[27];  // contribution to grid_v_in
  S m[27];       // contribution to grid_m
  int node[27];
};

// p2g body, mpm_simulator.py:198-225
template <class S, class T>
inline void p2g_body(const Consts<T>& k, const V3<S>& x, const V3<S>& v, const M3<S>& C, const M3<S>& Ftmp,
                     const M3<S>& U, const V3<S>& sig, const M3<S>& V, T mu, T lam, T ys, P2GOut<S, T>& o) {
  int base[3];
  V3<S> fx, w[3];
  bspline(k, x, base, fx, w);
  M3<S> nF = von_mises<S, T>(Ftmp, U, sig, V, ys, mu);
  o.newF = nF;
  S J = det3(nF);
  M3<S> r = matmul(U, transpose(V));
  M3<S> A;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A.a[i][j] = nF.a[i][j] - r.a[i][j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A.a[i][j] = (T(2) * mu) * A.a[i][j];
  M3<S> st = matmul(A, transpose(nF));
  for (int i = 0; i < 3; i++) st.a[i][i] = st.a[i][i] + (lam * J) * (J - T(1));
  M3<S> aff;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) aff.a[i][j] = k.c_stress * st.a[i][j] + k.p_mass * C.a[i][j];
  int q = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int l = 0; l < 3; l++, q++) {
        V3<S> dpos{{(T(i) - fx[0]) * k.dx, (T(j) - fx[1]) * k.dx, (T(l) - fx[2]) * k.dx}};
        S weight = T(1) * w[i][0];
        weight = weight * w[j][1];
        weight = weight * w[l][2];
        V3<S> ad = matvec(aff, dpos);
        for (int d = 0; d < 3; d++) o.mv[q][d] = weight * (k.p_mass * v[d] + ad[d]);
        o.m[q] = weight * k.p_mass;
        o.node[q] = ((base[0] + i) * k.n + (base[1] + j)) * k.n + (base[2] + l);
      }
}

// g2p body, mpm_simulator.py:264-283
template <class S, class T>
inline void g2p_body(const Consts<T>& k, const V3<S>& x, const V3<S> gv[27], V3<S>& nx, V3<S>& nv, M3<S>& nC) {
  int base[3];
  V3<S> fx, w[3];
  bspline(k, x, base, fx, w);
  for (int a = 0; a < 3; a++) {
    nv[a] = S(T(0));
    for (int b = 0; b < 3; b++) nC.a[a][b] = S(T(0));
  }
  int q = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int l = 0; l < 3; l++, q++) {
        V3<S> dpos{{T(i) - fx[0], T(j) - fx[1], T(l) - fx[2]}};
        S weight = T(1) * w[i][0];
        weight = weight * w[j][1];
        weight = weight * w[l][2];
        for (int a = 0; a < 3; a++) {
          nv[a] = nv[a] + weight * gv[q][a];
          for (int b = 0; b < 3; b++) nC.a[a][b] = nC.a[a][b] + (k.c_C * weight) * (gv[q][a] * dpos[b]);
        }
      }
  for (int d = 0; d < 3; d++) nx[d] = s_max(s_min(x[d] + k.dt * nv[d], k.x_hi), k.x_lo);
}

// grid_op body for one node, mpm_simulator.py:230-262
template <class S, class T>
inline V3<S> grid_op_body(const Consts<T>& k, const int I[3], const V3<S>& v_in, const S& m, int n_tools,
                          const ToolC<T>* tools, const Pose<S>* P0, const Pose<S>* P1) {
  S inv = T(1) / m;
  V3<S> v{{inv * v_in[0], inv * v_in[1], inv * v_in[2]}};
  for (int d = 0; d < 3; d++) v[d] = v[d] + (k.dt * k.gravity[d]) * T(30);
  V3<S> gp{{S(T(I[0]) * k.dx), S(T(I[1]) * k.dx), S(T(I[2]) * k.dx)}};
  for (int t = 0; t < n_tools; t++) v = tool_collide<S, T>(tools[t], P0[t], P1[t], gp, v, k.dt);
  const int bound = 3;
  for (int d = 0; d < 3; d++) {
    if (I[d] < bound && val(v[d]) < T(0)) {
      if (d != 1 || k.ground_friction_d == 0.0) {
        v[d] = S(T(0));
      } else if (k.ground_friction_d < 10.0) {
        S lin = v[1] + T(1e-30);  // v.dot(normal) + 1e-30, normal = e_y
        V3<S> vit;
        for (int a = 0; a < 3; a++) vit[a] = v[a] - lin * (a == 1 ? T(1) : T(0)) - T(I[a]) * T(1e-30);
        S lit = s_sqrt(dot(vit, vit) + T(1e-8));
        S sc = s_max(T(1) + k.ground_friction * lin / lit, T(0));
        for (int a = 0; a < 3; a++) v[a] = sc * (vit[a] + T(I[a]) * T(1e-30));
        v[1] = S(T(0));
      } else {
        for (int a = 0; a < 3; a++) v[a] = S(T(0));
      }
    }
    if (I[d] > k.n - bound && val(v[d]) > T(0)) v[d] = S(T(0));
  }
  return v;
}

// ----------------------------------------------------------------------------
// simulator state (fields as in MPMSimulator.__init__ / Primitive.__init__)
// ----------------------------------------------------------------------------
// ---- emulation of the reference's unspecified float-atomic order (test infrastructure) -------------------------------
// The reference accumulates p2g and the g2p adjoint with float atomics whose order changes from run to run, so every
// grid sum carries about an ulp of run-to-run noise.  With a non-zero seed the oracle multiplies each rounded grid sum
// by (1 + s * ulp), s in {-1, 0, +1} a hash of (seed, frame, node, component): the spread of the results over a few
// seeds is the reference formulation's own reproducibility floor for a scene (tests use it as the parity floor where it
// exceeds the north-star tolerance).  Deterministic in (seed, frame, node), so substep_grad's recompute sees the same grid.
static int g_scatter_noise_seed = 0;
static double g_scatter_noise_ulps = 1.0;
static int g_pose_atomics_seed = 0;   // != 0: tool-pose adjoints of grid_op.grad accumulate like the reference's atomics
static inline double scatter_noise(int f, int g, int d, double ulp) {
  if (g_scatter_noise_seed == 0) return 1.0;
  uint32_t h = (uint32_t)g_scatter_noise_seed * 0x9E3779B1u;
  h ^= (uint32_t)f + 0x7F4A7C15u + (h << 6) + (h >> 2);
  h ^= (uint32_t)g * 0x85EBCA6Bu + (h << 6) + (h >> 2);
  h ^= (uint32_t)d * 0xC2B2AE35u + (h << 6) + (h >> 2);
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return 1.0 + ((int)(h % 3u) - 1) * ulp * g_scatter_noise_ulps;
}
template <class T> inline double ulp_of();
template <> inline double ulp_of<float>() { return 1.1920928955078125e-7; }
template <> inline double ulp_of<double>() { return 2.220446049250313e-16; }

template <class T>
struct ToolState {
  ToolC<T> c;
  std::vector<T> pos, rot, vel, w, gap, gap_vel;          // [frames][..]
  std::vector<T> g_pos, g_rot, g_vel, g_w, g_gap, g_gap_vel;
  std::vector<T> action, g_action;                         // [frames][action_dim]
};

template <class T>
struct Sim {
  orc_config cfg;
  Consts<T> k;
  int n = 0, cap, frames, G, K, npairs;
  std::vector<T> x, v, C, F, gx, gv, gC, gF;
  std::vector<T> mu, lam, ys;
  std::vector<T> Ftmp, U, sig, V, gFtmp, gU, gsig, gV;
  std::vector<T> grid_v_in, grid_v_out, grid_m, g_grid_v_in, g_grid_v_out, g_grid_m;
  // The reference scatters with float atomics whose order is unspecified (GPU) -- the oracle accumulates every
  // scatter in double and rounds once: the order-independent centre of all valid fp32 results.
  std::vector<double> acc4, acc3;
  std::vector<ToolState<T>> tools;
  std::vector<T> rand_num, rand_points, g_rand_points;
  std::vector<int> collision_idx;
  std::vector<std::vector<T>> dists, g_dists;

  explicit Sim(const orc_config& c) : cfg(c) {
    cap = c.max_particles;
    frames = c.max_frames;
    K = c.n_tools;
    npairs = c.n_pairs;
    k.n = c.n_grid;
    G = k.n * k.n * k.n;
    k.dt = (T)c.dt;
    k.dx = (T)c.dx;
    k.inv_dx = (T)c.inv_dx;
    k.p_mass = (T)c.p_mass;
    k.c_stress = (T)(-c.dt * c.p_vol * 4 * c.inv_dx * c.inv_dx);  // mpm_simulator.py:214 (python double, then cast)
    k.c_C = (T)(4 * c.inv_dx);                                    // :279
    k.x_hi = (T)(1. - 3 * c.dx);                                  // :283
    k.x_lo = (T)(c.lower_bound * c.dx);
    k.m_eps = (T)1e-12;
    k.ground_friction = (T)c.ground_friction;
    k.ground_friction_d = c.ground_friction;
    for (int d = 0; d < 3; d++) k.gravity[d] = (T)c.gravity[d];
    size_t fp = (size_t)frames * cap;
    x.assign(fp * 3, 0);
    v.assign(fp * 3, 0);
    C.assign(fp * 9, 0);
    F.assign(fp * 9, 0);
    gx.assign(fp * 3, 0);
    gv.assign(fp * 3, 0);
    gC.assign(fp * 9, 0);
    gF.assign(fp * 9, 0);
    mu.assign(cap, (T)c.mu);
    lam.assign(cap, (T)c.lam);
    ys.assign(cap, (T)c.yield_stress);
    for (auto* p : {&Ftmp, &U, &V, &gFtmp, &gU, &gV}) p->assign((size_t)cap * 9, 0);
    sig.assign((size_t)cap * 3, 0);
    gsig.assign((size_t)cap * 3, 0);
    for (auto* p : {&grid_v_in, &grid_v_out, &g_grid_v_in, &g_grid_v_out}) p->assign((size_t)G * 3, 0);
    grid_m.assign(G, 0);
    g_grid_m.assign(G, 0);
    acc4.assign((size_t)G * 4, 0.0);
    acc3.assign((size_t)G * 3, 0.0);
    tools.resize(K);
    for (int i = 0; i < K; i++) {
      const orc_tool_cfg& tc = c.tools[i];
      ToolC<T>& cc = tools[i].c;
      cc.type = tc.type;
      cc.action_dim = tc.action_dim;
      for (int j = 0; j < 8; j++) cc.action_scale[j] = (T)tc.action_scale[j];
      cc.friction = (T)tc.friction;
      cc.softness = (T)tc.softness;
      for (int j = 0; j < 3; j++) {
        cc.lo[j] = (T)tc.lower_bound[j];
        cc.hi[j] = (T)tc.upper_bound[j];
        cc.size[j] = (T)tc.size[j];
      }
      cc.h = (T)tc.h;
      cc.half_h = (T)(tc.h / 2);
      cc.r = (T)tc.r;
      cc.radius = (T)tc.radius;
      cc.prism_h[0] = (T)tc.prism_h[0];
      cc.prism_h[1] = (T)tc.prism_h[1];
      for (int j = 0; j < 4; j++) cc.prot[j] = (T)tc.prot[j];
      cc.min_gap = (T)tc.minimal_gap;
      cc.max_gap = (T)tc.maximal_gap;
      ToolState<T>& ts = tools[i];
      ts.pos.assign((size_t)frames * 3, 0);
      ts.rot.assign((size_t)frames * 4, 0);
      ts.vel.assign((size_t)frames * 3, 0);
      ts.w.assign((size_t)frames * 3, 0);
      ts.gap.assign(frames, 0);
      ts.gap_vel.assign(frames, 0);
      ts.g_pos = ts.pos;
      ts.g_rot = ts.rot;
      ts.g_vel = ts.vel;
      ts.g_w = ts.w;
      ts.g_gap = ts.gap;
      ts.g_gap_vel = ts.gap_vel;
      int ad = std::max(1, cc.action_dim);
      ts.action.assign((size_t)frames * ad, 0);
      ts.g_action.assign((size_t)frames * ad, 0);
    }
    if (npairs > 0) {
      rand_num.assign((size_t)npairs * ORC_NUM_COLLISION_POINTS * 3, 0);
      rand_points.assign((size_t)frames * npairs * ORC_NUM_COLLISION_POINTS * 3, 0);
      g_rand_points = rand_points;
      collision_idx.assign((size_t)frames * npairs, -1);
    }
    int ncols = 0;
    for (int i = 0; i < K; i++) ncols += is_gripper(tools[i].c.type) ? 2 : 1;
    dists.assign(ncols, std::vector<T>(cap, 0));
    g_dists = dists;
  }

  // ---- helpers -------------------------------------------------------------
  template <class S>
  Pose<S> load_pose(int i, int f, bool as_input) const {
    const ToolState<T>& t = tools[i];
    Pose<S> P;
    for (int d = 0; d < 3; d++) P.pos[d] = mkS<S>(t.pos[f * 3 + d], as_input);
    for (int d = 0; d < 4; d++) P.rot[d] = mkS<S>(t.rot[f * 4 + d], as_input);
    P.gap = mkS<S>(t.gap[f], as_input);
    return P;
  }
  template <class S>
  static S mkS(T c, bool as_input) {
    return mkS_impl(c, as_input, (S*)nullptr);
  }
  static T mkS_impl(T c, bool, T*) { return c; }
  static Var<T> mkS_impl(T c, bool as_input, Var<T>*) { return as_input ? Var<T>::input(c) : Var<T>(c); }

  void add_pose_grad(int i, int f, const Pose<Var<T>>& P) {
    ToolState<T>& t = tools[i];
    for (int d = 0; d < 3; d++) t.g_pos[f * 3 + d] += P.pos[d].grad();
    for (int d = 0; d < 4; d++) t.g_rot[f * 4 + d] += P.rot[d].grad();
    t.g_gap[f] += P.gap.grad();
  }

  // ---- forward kernels ------------------------------------------------------
  void clear_grid() {  // mpm_simulator.py:100-110
    std::fill(grid_v_in.begin(), grid_v_in.end(), T(0));
    std::fill(grid_v_out.begin(), grid_v_out.end(), T(0));
    std::fill(grid_m.begin(), grid_m.end(), T(0));
    std::fill(g_grid_v_in.begin(), g_grid_v_in.end(), T(0));
    std::fill(g_grid_v_out.begin(), g_grid_v_out.end(), T(0));
    std::fill(g_grid_m.begin(), g_grid_m.end(), T(0));
  }
  void clear_SVD_grad() {  // :112-119
    std::fill(gU.begin(), gU.end(), T(0));
    std::fill(gsig.begin(), gsig.end(), T(0));
    std::fill(gV.begin(), gV.end(), T(0));
    std::fill(gFtmp.begin(), gFtmp.end(), T(0));
  }
  template <class S>
  static M3<S> ftmp_body(T dt, const M3<S>& C, const M3<S>& F) {  // :121-124
    M3<S> M;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) M.a[i][j] = (i == j ? T(1) : T(0)) + dt * C.a[i][j];
    return matmul(M, F);
  }
  template <class S>
  M3<S> load_m3(const std::vector<T>& a, size_t idx, bool in) const {
    M3<S> m;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) m.a[i][j] = mkS<S>(a[idx * 9 + i * 3 + j], in);
    return m;
  }
  template <class S>
  V3<S> load_v3(const std::vector<T>& a, size_t idx, bool in) const {
    V3<S> m;
    for (int i = 0; i < 3; i++) m[i] = mkS<S>(a[idx * 3 + i], in);
    return m;
  }
  void compute_F_tmp(int f) {
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      size_t fp = (size_t)f * cap + p;
      M3<T> r = ftmp_body<T>(k.dt, load_m3<T>(C, fp, false), load_m3<T>(F, fp, false));
      for (int i = 0; i < 9; i++) Ftmp[(size_t)p * 9 + i] = r.a[i / 3][i % 3];
    }
  }
  void svd() {  // :126-129
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      if (g_svd_algorithm == 1) svd3_eig_qr<T>(&Ftmp[(size_t)p * 9], &U[(size_t)p * 9], &sig[(size_t)p * 3], &V[(size_t)p * 9]);
      else svd3<T>(&Ftmp[(size_t)p * 9], &U[(size_t)p * 9], &sig[(size_t)p * 3], &V[(size_t)p * 9]);
    }
  }
  void p2g(int f) {
    std::fill(acc4.begin(), acc4.end(), 0.0);
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      size_t fp = (size_t)f * cap + p;
      P2GOut<T, T> o;
      p2g_body<T, T>(k, load_v3<T>(x, fp, false), load_v3<T>(v, fp, false), load_m3<T>(C, fp, false),
                     load_m3<T>(Ftmp, p, false), load_m3<T>(U, p, false), load_v3<T>(sig, p, false),
                     load_m3<T>(V, p, false), mu[p], lam[p], ys[p], o);
      size_t fp1 = (size_t)(f + 1) * cap + p;
      for (int i = 0; i < 9; i++) F[fp1 * 9 + i] = o.newF.a[i / 3][i % 3];
      for (int q = 0; q < 27; q++) {
        for (int d = 0; d < 3; d++) {
#pragma omp atomic
          acc4[(size_t)o.node[q] * 4 + d] += (double)o.mv[q][d];
        }
#pragma omp atomic
        acc4[(size_t)o.node[q] * 4 + 3] += (double)o.m[q];
      }
    }
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; g++) {
      for (int d = 0; d < 3; d++)
        grid_v_in[(size_t)g * 3 + d] = (T)((double)(T)acc4[(size_t)g * 4 + d] * scatter_noise(f, g, d, ulp_of<T>()));
      grid_m[g] = (T)((double)(T)acc4[(size_t)g * 4 + 3] * scatter_noise(f, g, 3, ulp_of<T>()));
    }
  }
  void forward_kinematics(int i, int f) {
    ToolState<T>& t = tools[i];
    Pose<T> P = load_pose<T>(i, f, false);
    V3<T> vv{{t.vel[f * 3], t.vel[f * 3 + 1], t.vel[f * 3 + 2]}}, ww{{t.w[f * 3], t.w[f * 3 + 1], t.w[f * 3 + 2]}};
    Pose<T> N = tool_fk<T, T>(t.c, P, vv, ww, t.gap_vel[f]);
    for (int d = 0; d < 3; d++) t.pos[(f + 1) * 3 + d] = N.pos[d];
    for (int d = 0; d < 4; d++) t.rot[(f + 1) * 4 + d] = N.rot[d];
    t.gap[f + 1] = N.gap;
  }
  template <class S>
  V3<S> surface_pos(int j, const Pose<S>& Pj, const T* rn) const {  // mpm_simulator.py:291-292, primitives.py:369-372
    const ToolC<T>& c = tools[j].c;
    V3<S> q;
    for (int d = 0; d < 3; d++) {
      T rp = rn[d] * c.size[d];
      q[d] = S(s_max(s_min(rp, c.size[d]), -c.size[d]));
    }
    return qrot(Pj.rot, q) + Pj.pos;
  }
  size_t rp_index(int f, int cid, int kk) const { return (((size_t)f * npairs + cid) * ORC_NUM_COLLISION_POINTS + kk) * 3; }
  void set_surface_points(int s) {  // :286-292
    for (int cid = 0; cid < npairs; cid++) {
      int j = cfg.pairs[cid][1];
      Pose<T> Pj = load_pose<T>(j, s + 1, false);
      for (int kk = 0; kk < ORC_NUM_COLLISION_POINTS; kk++) {
        V3<T> r = surface_pos<T>(j, Pj, &rand_num[((size_t)cid * ORC_NUM_COLLISION_POINTS + kk) * 3]);
        for (int d = 0; d < 3; d++) rand_points[rp_index(s + 1, cid, kk) + d] = r[d];
      }
    }
  }
  void set_collision_idx(int s) {  // :294-298 ; primive_base.py:135-143 made deterministic (first minimum)
    for (int cid = 0; cid < npairs; cid++) {
      int i = cfg.pairs[cid][0];
      Pose<T> Pi = load_pose<T>(i, s + 1, false);
      T min_dist = 0;
      int idx = -1;
      for (int kk = 0; kk < ORC_NUM_COLLISION_POINTS; kk++) {
        size_t o = rp_index(s + 1, cid, kk);
        V3<T> p{{rand_points[o], rand_points[o + 1], rand_points[o + 2]}};
        T dist = tool_sdf<T, T>(tools[i].c, Pi, p);
        if (dist < min_dist) {
          min_dist = dist;
          idx = kk;
        }
      }
      collision_idx[(size_t)(s + 1) * npairs + cid] = idx;
    }
  }
  template <class S>
  V3<S> projection_body(int i, const Pose<S>& Pi, const V3<S>& pt) const {  // primive_base.py:145-150
    typedef T TT;
    S dist = tool_sdf<S, T>(tools[i].c, Pi, pt);
    V3<S> nr = tool_normal<S, T>(tools[i].c, Pi, pt);
    S inv = TT(1) / s_sqrt(dot(nr, nr));
    V3<S> r;
    for (int d = 0; d < 3; d++) r[d] = Pi.pos[d] + (inv * nr[d]) * dist;
    return r;
  }
  void apply_collision_projection(int s) {  // :300-305
    for (int cid = 0; cid < npairs; cid++) {
      int i = cfg.pairs[cid][0];
      int idx = collision_idx[(size_t)(s + 1) * npairs + cid];
      if (idx != -1) {
        size_t o = rp_index(s + 1, cid, idx);
        V3<T> p{{rand_points[o], rand_points[o + 1], rand_points[o + 2]}};
        V3<T> np = projection_body<T>(i, load_pose<T>(i, s + 1, false), p);
        for (int d = 0; d < 3; d++) tools[i].pos[(s + 1) * 3 + d] = np[d];
      }
    }
  }
  void grid_op(int f) {
    std::vector<Pose<T>> P0(K), P1(K);
    std::vector<ToolC<T>> tc(K);
    for (int i = 0; i < K; i++) {
      P0[i] = load_pose<T>(i, f, false);
      P1[i] = load_pose<T>(i, f + 1, false);
      tc[i] = tools[i].c;
    }
    int nn = k.n;
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; g++) {
      if (grid_m[g] > k.m_eps) {
        int I[3] = {g / (nn * nn), (g / nn) % nn, g % nn};
        V3<T> vin{{grid_v_in[(size_t)g * 3], grid_v_in[(size_t)g * 3 + 1], grid_v_in[(size_t)g * 3 + 2]}};
        V3<T> vo = grid_op_body<T, T>(k, I, vin, grid_m[g], K, tc.data(), P0.data(), P1.data());
        for (int d = 0; d < 3; d++) grid_v_out[(size_t)g * 3 + d] = vo[d];
      }
    }
  }
  void g2p(int f) {
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      size_t fp = (size_t)f * cap + p, fp1 = (size_t)(f + 1) * cap + p;
      V3<T> xx = load_v3<T>(x, fp, false);
      int base[3];
      V3<T> fx, w[3];
      bspline(k, xx, base, fx, w);
      V3<T> gvv[27];
      int q = 0;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          for (int l = 0; l < 3; l++, q++) {
            size_t node = ((size_t)(base[0] + i) * k.n + (base[1] + j)) * k.n + (base[2] + l);
            for (int d = 0; d < 3; d++) gvv[q][d] = grid_v_out[node * 3 + d];
          }
      V3<T> nx, nv;
      M3<T> nC;
      g2p_body<T, T>(k, xx, gvv, nx, nv, nC);
      for (int d = 0; d < 3; d++) {
        v[fp1 * 3 + d] = nv[d];
        x[fp1 * 3 + d] = nx[d];
      }
      for (int i = 0; i < 9; i++) C[fp1 * 9 + i] = nC.a[i / 3][i % 3];
    }
  }
  void substep(int s) {  // :307-323
    clear_grid();
    compute_F_tmp(s);
    svd();
    p2g(s);
    for (int i = 0; i < K; i++) forward_kinematics(i, s);
    if (npairs > 0) {
      set_surface_points(s);
      set_collision_idx(s);
      apply_collision_projection(s);
    }
    grid_op(s);
    g2p(s);
  }

  // ---- adjoint kernels ------------------------------------------------------
  void g2p_grad(int f) {
    std::fill(acc3.begin(), acc3.end(), 0.0);
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      typedef Var<T> S;
      Tape<T>& tp = tape<T>();
      tp.clear();
      size_t fp = (size_t)f * cap + p, fp1 = (size_t)(f + 1) * cap + p;
      V3<S> xx = load_v3<S>(x, fp, true);
      int base[3];
      {
        V3<T> xt = load_v3<T>(x, fp, false), fx, w[3];
        bspline(k, xt, base, fx, w);
      }
      V3<S> gvv[27];
      size_t nodes[27];
      int q = 0;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          for (int l = 0; l < 3; l++, q++) {
            nodes[q] = ((size_t)(base[0] + i) * k.n + (base[1] + j)) * k.n + (base[2] + l);
            for (int d = 0; d < 3; d++) gvv[q][d] = S::input(grid_v_out[nodes[q] * 3 + d]);
          }
      V3<S> nx, nv;
      M3<S> nC;
      g2p_body<S, T>(k, xx, gvv, nx, nv, nC);
      tp.begin_reverse();
      for (int d = 0; d < 3; d++) {
        nx[d].seed(gx[fp1 * 3 + d]);
        nv[d].seed(gv[fp1 * 3 + d]);
      }
      for (int i = 0; i < 9; i++) nC.a[i / 3][i % 3].seed(gC[fp1 * 9 + i]);
      tp.reverse();
      for (int d = 0; d < 3; d++) gx[fp * 3 + d] += xx[d].grad();
      for (q = 0; q < 27; q++)
        for (int d = 0; d < 3; d++) {
          double g = (double)gvv[q][d].grad();
#pragma omp atomic
          acc3[nodes[q] * 3 + d] += g;
        }
    }
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; g++)
      for (int d = 0; d < 3; d++)
        g_grid_v_out[(size_t)g * 3 + d] += (T)((double)(T)acc3[(size_t)g * 3 + d] * scatter_noise(f, g, 4 + d, ulp_of<T>()));
  }
  void grid_op_grad(int f) {
    int nn = k.n;
    std::vector<ToolC<T>> tc(K);
    for (int i = 0; i < K; i++) tc[i] = tools[i].c;
    std::vector<int> live;
    for (int g = 0; g < G; g++)
      if (grid_m[g] > k.m_eps) live.push_back(g);
    const size_t W = (size_t)K * 16;
    std::vector<T> contrib(live.size() * W, T(0));   // per-node adjoints of (pose f, pose f+1) of every tool
#pragma omp parallel
    {
      typedef Var<T> S;
      std::vector<Pose<S>> P0(K), P1(K);
#pragma omp for schedule(static)
      for (int li = 0; li < (int)live.size(); li++) {
        int g = live[li];
        Tape<T>& tp = tape<T>();
        tp.clear();
        int I[3] = {g / (nn * nn), (g / nn) % nn, g % nn};
        V3<S> vin{{S::input(grid_v_in[(size_t)g * 3]), S::input(grid_v_in[(size_t)g * 3 + 1]),
                   S::input(grid_v_in[(size_t)g * 3 + 2])}};
        S m = S::input(grid_m[g]);
        for (int i = 0; i < K; i++) {
          P0[i] = load_pose<S>(i, f, true);
          P1[i] = load_pose<S>(i, f + 1, true);
        }
        V3<S> vo = grid_op_body<S, T>(k, I, vin, m, K, tc.data(), P0.data(), P1.data());
        tp.begin_reverse();
        for (int d = 0; d < 3; d++) vo[d].seed(g_grid_v_out[(size_t)g * 3 + d]);
        tp.reverse();
        for (int d = 0; d < 3; d++) g_grid_v_in[(size_t)g * 3 + d] += vin[d].grad();
        g_grid_m[g] += m.grad();
        for (int i = 0; i < K; i++) {
          T* a = &contrib[(size_t)li * W + (size_t)i * 16];
          for (int d = 0; d < 3; d++) a[d] = P0[i].pos[d].grad();
          for (int d = 0; d < 4; d++) a[3 + d] = P0[i].rot[d].grad();
          a[7] = P0[i].gap.grad();
          for (int d = 0; d < 3; d++) a[8 + d] = P1[i].pos[d].grad();
          for (int d = 0; d < 4; d++) a[11 + d] = P1[i].rot[d].grad();
          a[15] = P1[i].gap.grad();
        }
      }
    }
    auto field = [&](ToolState<T>& ts, int q) -> T& {   // slot q of (pose f | pose f+1) in the tool's .grad fields
      int ff = q < 8 ? f : f + 1, c = q & 7;
      return c < 3 ? ts.g_pos[ff * 3 + c] : (c < 7 ? ts.g_rot[ff * 4 + c - 3] : ts.g_gap[ff]);
    };
    if (g_pose_atomics_seed != 0) {
      // the reference: every node thread atomically adds its term to position.grad / rotation.grad in T precision, in an
      // unspecified order (the terms are O(1/dt) and cancel, so the order matters): seeded shuffle of the node order
      std::vector<int> order(live.size());
      for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
      uint32_t st = (uint32_t)g_pose_atomics_seed * 2654435761u + (uint32_t)f * 40503u + 1u;
      for (size_t i = order.size(); i > 1; i--) {
        st ^= st << 13; st ^= st >> 17; st ^= st << 5;
        std::swap(order[i - 1], order[st % i]);
      }
      for (int li : order)
        for (int i = 0; i < K; i++)
          for (int q = 0; q < 16; q++) {
            T& dst = field(tools[i], q);
            dst = dst + contrib[(size_t)li * W + (size_t)i * 16 + q];
          }
      return;
    }
    // default: order-independent (sum in double, round once)
    for (int i = 0; i < K; i++)
      for (int q = 0; q < 16; q++) {
        double a = 0.0;
        for (size_t li = 0; li < live.size(); li++) a += (double)contrib[li * W + (size_t)i * 16 + q];
        field(tools[i], q) += (T)a;
      }
  }
  void apply_collision_projection_grad(int s) {
    typedef Var<T> S;
    for (int cid = npairs - 1; cid >= 0; cid--) {
      int i = cfg.pairs[cid][0];
      int idx = collision_idx[(size_t)(s + 1) * npairs + cid];
      if (idx == -1) continue;
      Tape<T>& tp = tape<T>();
      tp.clear();
      size_t o = rp_index(s + 1, cid, idx);
      V3<S> pt{{S::input(rand_points[o]), S::input(rand_points[o + 1]), S::input(rand_points[o + 2])}};
      // Taichi re-evaluates the forward at the CURRENT (post-projection) field value
      Pose<S> Pi = load_pose<S>(i, s + 1, true);
      V3<S> np = projection_body<S>(i, Pi, pt);
      tp.begin_reverse();
      ToolState<T>& ts = tools[i];
      for (int d = 0; d < 3; d++) np[d].seed(ts.g_pos[(s + 1) * 3 + d]);
      tp.reverse();
      // position.grad keeps its value (y += shift) and additionally receives d(shift)/d(position)
      for (int d = 0; d < 3; d++) ts.g_pos[(s + 1) * 3 + d] = Pi.pos[d].grad();
      for (int d = 0; d < 4; d++) ts.g_rot[(s + 1) * 4 + d] += Pi.rot[d].grad();
      ts.g_gap[s + 1] += Pi.gap.grad();
      for (int d = 0; d < 3; d++) g_rand_points[o + d] += pt[d].grad();
    }
  }
  void set_surface_points_grad(int s) {
    typedef Var<T> S;
    for (int cid = npairs - 1; cid >= 0; cid--) {
      int j = cfg.pairs[cid][1];
      for (int kk = 0; kk < ORC_NUM_COLLISION_POINTS; kk++) {
        size_t o = rp_index(s + 1, cid, kk);
        if (g_rand_points[o] == T(0) && g_rand_points[o + 1] == T(0) && g_rand_points[o + 2] == T(0)) continue;
        Tape<T>& tp = tape<T>();
        tp.clear();
        Pose<S> Pj = load_pose<S>(j, s + 1, true);
        V3<S> r = surface_pos<S>(j, Pj, &rand_num[((size_t)cid * ORC_NUM_COLLISION_POINTS + kk) * 3]);
        tp.begin_reverse();
        for (int d = 0; d < 3; d++) r[d].seed(g_rand_points[o + d]);
        tp.reverse();
        add_pose_grad(j, s + 1, Pj);
      }
    }
  }
  void forward_kinematics_grad(int i, int f) {
    typedef Var<T> S;
    Tape<T>& tp = tape<T>();
    tp.clear();
    ToolState<T>& t = tools[i];
    Pose<S> P = load_pose<S>(i, f, true);
    V3<S> vv{{S::input(t.vel[f * 3]), S::input(t.vel[f * 3 + 1]), S::input(t.vel[f * 3 + 2])}};
    V3<S> ww{{S::input(t.w[f * 3]), S::input(t.w[f * 3 + 1]), S::input(t.w[f * 3 + 2])}};
    S gvl = S::input(t.gap_vel[f]);
    Pose<S> N = tool_fk<S, T>(t.c, P, vv, ww, gvl);
    tp.begin_reverse();
    for (int d = 0; d < 3; d++) N.pos[d].seed(t.g_pos[(f + 1) * 3 + d]);
    for (int d = 0; d < 4; d++) N.rot[d].seed(t.g_rot[(f + 1) * 4 + d]);
    N.gap.seed(t.g_gap[f + 1]);
    tp.reverse();
    add_pose_grad(i, f, P);
    for (int d = 0; d < 3; d++) {
      t.g_vel[f * 3 + d] += vv[d].grad();
      t.g_w[f * 3 + d] += ww[d].grad();
    }
    t.g_gap_vel[f] += gvl.grad();
  }
  void p2g_grad(int f) {
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      typedef Var<T> S;
      Tape<T>& tp = tape<T>();
      tp.clear();
      size_t fp = (size_t)f * cap + p, fp1 = (size_t)(f + 1) * cap + p;
      V3<S> xx = load_v3<S>(x, fp, true), vv = load_v3<S>(v, fp, true);
      M3<S> CC = load_m3<S>(C, fp, true), Ft = load_m3<S>(Ftmp, p, true), UU = load_m3<S>(U, p, true),
            VV = load_m3<S>(V, p, true);
      V3<S> ss = load_v3<S>(sig, p, true);
      P2GOut<S, T> o;
      p2g_body<S, T>(k, xx, vv, CC, Ft, UU, ss, VV, mu[p], lam[p], ys[p], o);
      tp.begin_reverse();
      for (int i = 0; i < 9; i++) o.newF.a[i / 3][i % 3].seed(gF[fp1 * 9 + i]);
      for (int q = 0; q < 27; q++) {
        for (int d = 0; d < 3; d++) o.mv[q][d].seed(g_grid_v_in[(size_t)o.node[q] * 3 + d]);
        o.m[q].seed(g_grid_m[o.node[q]]);
      }
      tp.reverse();
      for (int d = 0; d < 3; d++) {
        gx[fp * 3 + d] += xx[d].grad();
        gv[fp * 3 + d] += vv[d].grad();
        gsig[(size_t)p * 3 + d] += ss[d].grad();
      }
      for (int i = 0; i < 9; i++) {
        gC[fp * 9 + i] += CC.a[i / 3][i % 3].grad();
        gFtmp[(size_t)p * 9 + i] += Ft.a[i / 3][i % 3].grad();
        gU[(size_t)p * 9 + i] += UU.a[i / 3][i % 3].grad();
        gV[(size_t)p * 9 + i] += VV.a[i / 3][i % 3].grad();
      }
    }
  }
  static T clampK(T a) {  // mpm_simulator.py:184-192
    if (a >= 0)
      a = std::max(a, (T)1e-6);
    else
      a = std::min(a, (T)-1e-6);
    return a;
  }
  void svd_grad() {  // mpm_simulator.py:131-156 (hand-written in the reference)
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      M3<T> u = load_m3<T>(U, p, false), vm = load_m3<T>(V, p, false), gu = load_m3<T>(gU, p, false),
            gvm = load_m3<T>(gV, p, false);
      M3<T> sg, gs;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          sg.a[i][j] = i == j ? sig[(size_t)p * 3 + i] : T(0);
          gs.a[i][j] = i == j ? gsig[(size_t)p * 3 + i] : T(0);
        }
      M3<T> vt = transpose(vm), ut = transpose(u);
      M3<T> sigma_term = matmul(matmul(u, gs), vt);
      T s2[3];
      for (int i = 0; i < 3; i++) s2[i] = sg.a[i][i] * sg.a[i][i];
      M3<T> Fm;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Fm.a[i][j] = i == j ? T(0) : T(1) / clampK(s2[j] - s2[i]);
      M3<T> a1 = matmul(ut, gu), a2 = matmul(transpose(gu), u), b1 = matmul(vt, gvm), b2 = matmul(transpose(gvm), vm);
      M3<T> A, B;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          A.a[i][j] = Fm.a[i][j] * (a1.a[i][j] - a2.a[i][j]);
          B.a[i][j] = Fm.a[i][j] * (b1.a[i][j] - b2.a[i][j]);
        }
      M3<T> u_term = matmul(matmul(u, matmul(A, sg)), vt);
      M3<T> v_term = matmul(u, matmul(sg, matmul(B, vt)));
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          gFtmp[(size_t)p * 9 + i * 3 + j] += (u_term.a[i][j] + v_term.a[i][j]) + sigma_term.a[i][j];
    }
  }
  void compute_F_tmp_grad(int f) {
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; p++) {
      typedef Var<T> S;
      Tape<T>& tp = tape<T>();
      tp.clear();
      size_t fp = (size_t)f * cap + p;
      M3<S> CC = load_m3<S>(C, fp, true), FF = load_m3<S>(F, fp, true);
      M3<S> r = ftmp_body<S>(k.dt, CC, FF);
      tp.begin_reverse();
      for (int i = 0; i < 9; i++) r.a[i / 3][i % 3].seed(gFtmp[(size_t)p * 9 + i]);
      tp.reverse();
      for (int i = 0; i < 9; i++) {
        gC[fp * 9 + i] += CC.a[i / 3][i % 3].grad();
        gF[fp * 9 + i] += FF.a[i / 3][i % 3].grad();
      }
    }
  }
  void substep_grad(int s) {  // :325-345
    clear_grid();
    clear_SVD_grad();
    compute_F_tmp(s);
    svd();
    p2g(s);
    grid_op(s);
    g2p_grad(s);
    grid_op_grad(s);
    if (npairs > 0) {
      apply_collision_projection_grad(s);
      set_surface_points_grad(s);
    }
    for (int i = K - 1; i >= 0; i--) forward_kinematics_grad(i, s);
    p2g_grad(s);
    svd_grad();
    compute_F_tmp_grad(s);
  }

  // ---- actions (primive_base.py:241-282, primitives.py:462-469, 863-867) ----
  void set_action(int s, int nsub, const double* action) {
    int off = 0;
    for (int i = 0; i < K; i++) {
      ToolState<T>& t = tools[i];
      int ad = t.c.action_dim;
      if (ad > 0) {
        for (int j = 0; j < ad; j++) {
          double a = std::min(1.0, std::max(-1.0, action[off + j]));
          t.action[(size_t)s * ad + j] = (T)a;
        }
        set_velocity(i, s, nsub);
      }
      off += ad;
    }
  }
  void set_velocity(int i, int s, int nsub) {
    ToolState<T>& t = tools[i];
    int ad = t.c.action_dim;
    const T* a = &t.action[(size_t)s * ad];
    for (int j = s * nsub; j < (s + 1) * nsub; j++) {
      for (int d = 0; d < 3; d++) t.vel[j * 3 + d] = a[d] * t.c.action_scale[d] / T(nsub);
      if (ad > 3)
        for (int d = 0; d < 3; d++) t.w[j * 3 + d] = a[d + 3] * t.c.action_scale[d + 3] / T(nsub);
      if (has_gap(t.c.type)) t.gap_vel[j] = a[6] * t.c.action_scale[6] / T(nsub);
    }
  }
  void set_velocity_grad(int s, int nsub) {
    for (int i = 0; i < K; i++) {
      ToolState<T>& t = tools[i];
      int ad = t.c.action_dim;
      if (ad <= 0) continue;
      T* ga = &t.g_action[(size_t)s * ad];
      for (int j = s * nsub; j < (s + 1) * nsub; j++) {
        for (int d = 0; d < 3; d++) ga[d] += t.g_vel[j * 3 + d] * (t.c.action_scale[d] / T(nsub));
        if (ad > 3)
          for (int d = 0; d < 3; d++) ga[d + 3] += t.g_w[j * 3 + d] * (t.c.action_scale[d + 3] / T(nsub));
        if (has_gap(t.c.type)) ga[6] += t.g_gap_vel[j] * (t.c.action_scale[6] / T(nsub));
      }
    }
  }
  void zero_grad() {
    for (auto* p : {&gx, &gv, &gC, &gF, &g_rand_points}) std::fill(p->begin(), p->end(), T(0));
    for (auto& t : tools)
      for (auto* p : {&t.g_pos, &t.g_rot, &t.g_vel, &t.g_w, &t.g_gap, &t.g_gap_vel, &t.g_action})
        std::fill(p->begin(), p->end(), T(0));
    for (auto& d : g_dists) std::fill(d.begin(), d.end(), T(0));
  }

  // ---- observation helpers --------------------------------------------------
  void compute_min_dist(int f) {  // function.py:79-88
    int col = 0;
    for (int j = 0; j < K; j++) {
      Pose<T> P = load_pose<T>(j, f, false);
      for (int p = 0; p < n; p++) {
        V3<T> xx = load_v3<T>(x, (size_t)f * cap + p, false);
        if (!is_gripper(tools[j].c.type)) {
          dists[col][p] = tool_sdf<T, T>(tools[j].c, P, xx);
        } else {
          dists[col][p] = sdf2<T, T>(tools[j].c, P, xx, -1);
          dists[col + 1][p] = sdf2<T, T>(tools[j].c, P, xx, 1);
        }
      }
      col += is_gripper(tools[j].c.type) ? 2 : 1;
    }
  }
  void compute_min_dist_grad(int f) {
    typedef Var<T> S;
    int col = 0;
    for (int j = 0; j < K; j++) {
      bool grip = is_gripper(tools[j].c.type);
      for (int p = 0; p < n; p++) {
        Tape<T>& tp = tape<T>();
        tp.clear();
        size_t fp = (size_t)f * cap + p;
        Pose<S> P = load_pose<S>(j, f, true);
        V3<S> xx = load_v3<S>(x, fp, true);
        S d0, d1;
        if (!grip) {
          d0 = tool_sdf<S, T>(tools[j].c, P, xx);
        } else {
          d0 = sdf2<S, T>(tools[j].c, P, xx, -1);
          d1 = sdf2<S, T>(tools[j].c, P, xx, 1);
        }
        tp.begin_reverse();
        d0.seed(g_dists[col][p]);
        if (grip) d1.seed(g_dists[col + 1][p]);
        tp.reverse();
        for (int d = 0; d < 3; d++) gx[fp * 3 + d] += xx[d].grad();
        add_pose_grad(j, f, P);
      }
      col += grip ? 2 : 1;
    }
  }
  template <class S>
  void grid_m_body(const V3<S>& xx, S out[27], size_t nodes[27]) const {  // mpm_simulator.py:456-466
    int base[3];
    V3<S> fx, w[3];
    bspline(k, xx, base, fx, w);
    int q = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        for (int l = 0; l < 3; l++, q++) {
          S weight = T(1) * w[i][0];
          weight = weight * w[j][1];
          weight = weight * w[l][2];
          out[q] = weight * k.p_mass;
          nodes[q] = ((size_t)(base[0] + i) * k.n + (base[1] + j)) * k.n + (base[2] + l);
        }
  }
  void compute_grid_m(int f) {
    std::fill(acc4.begin(), acc4.end(), 0.0);
    for (int p = 0; p < n; p++) {
      T o[27];
      size_t nodes[27];
      grid_m_body<T>(load_v3<T>(x, (size_t)f * cap + p, false), o, nodes);
      for (int q = 0; q < 27; q++) acc4[nodes[q] * 4 + 3] += (double)o[q];
    }
    for (int g = 0; g < G; g++) grid_m[g] = (T)acc4[(size_t)g * 4 + 3];
  }
  void compute_grid_m_grad(int f) {
    typedef Var<T> S;
    for (int p = 0; p < n; p++) {
      Tape<T>& tp = tape<T>();
      tp.clear();
      size_t fp = (size_t)f * cap + p;
      V3<S> xx = load_v3<S>(x, fp, true);
      S o[27];
      size_t nodes[27];
      grid_m_body<S>(xx, o, nodes);
      tp.begin_reverse();
      for (int q = 0; q < 27; q++) o[q].seed(g_grid_m[nodes[q]]);
      tp.reverse();
      for (int d = 0; d < 3; d++) gx[fp * 3 + d] += xx[d].grad();
    }
  }
};

struct Handle {
  int f64;
  Sim<float>* f;
  Sim<double>* d;
};
}  // namespace orc

using namespace orc;
#define DISPATCH(h, ...)             \
  do {                               \
    Handle* H_ = (Handle*)(h);       \
    if (H_->f64) {                   \
      auto& S = *H_->d;              \
      __VA_ARGS__;                   \
    } else {                         \
      auto& S = *H_->f;              \
      __VA_ARGS__;                   \
    }                                \
  } while (0)

template <class T>
static void copy_in(std::vector<T>& dst, size_t off, const double* src, size_t cnt) {
  for (size_t i = 0; i < cnt; i++) dst[off + i] = (T)src[i];
}
template <class T>
static void copy_out(const std::vector<T>& src, size_t off, double* dst, size_t cnt) {
  for (size_t i = 0; i < cnt; i++) dst[i] = (double)src[off + i];
}
template <class T>
static void add_in(std::vector<T>& dst, size_t off, const double* src, size_t cnt) {
  for (size_t i = 0; i < cnt; i++) dst[off + i] += (T)src[i];
}

/* adjoint probes (tape AD of the same templates the .grad kernels use): out = [g(p) 3 | g(v_in) 3 | g(pose f) 8 | g(pose f+1) 8] */
static const int PROBE_OUT = 22;
template <class SimT>
static void probe_grad(SimT& S, int tool, int f, int what, const double* p, const double* v_in, const double* gout,
                       double* out) {
  typedef decltype(S.k.dt) TT;
  typedef Var<TT> VS;
  Tape<TT>& tp = tape<TT>();
  tp.clear();
  V3<VS> pp{{VS::input((TT)p[0]), VS::input((TT)p[1]), VS::input((TT)p[2])}};
  V3<VS> vv{{VS::input((TT)v_in[0]), VS::input((TT)v_in[1]), VS::input((TT)v_in[2])}};
  Pose<VS> P0 = S.template load_pose<VS>(tool, f, true), P1 = S.template load_pose<VS>(tool, f + 1, true);
  V3<VS> r3;
  VS r1;
  int nout = 3;
  if (what == 0) {
    r1 = tool_sdf<VS, TT>(S.tools[tool].c, P0, pp);
    nout = 1;
  } else if (what == 1) {
    r3 = tool_normal<VS, TT>(S.tools[tool].c, P0, pp);
  } else {
    r3 = tool_collide<VS, TT>(S.tools[tool].c, P0, P1, pp, vv, S.k.dt);
  }
  tp.begin_reverse();
  if (nout == 1) r1.seed((TT)gout[0]);
  else for (int d = 0; d < 3; d++) r3[d].seed((TT)gout[d]);
  tp.reverse();
  for (int d = 0; d < 3; d++) out[d] = (double)pp[d].grad();
  for (int d = 0; d < 3; d++) out[3 + d] = (double)vv[d].grad();
  const Pose<VS>* Ps[2] = {&P0, &P1};
  for (int q = 0; q < 2; q++) {
    for (int d = 0; d < 3; d++) out[6 + q * 8 + d] = (double)Ps[q]->pos[d].grad();
    for (int d = 0; d < 4; d++) out[6 + q * 8 + 3 + d] = (double)Ps[q]->rot[d].grad();
    out[6 + q * 8 + 7] = (double)Ps[q]->gap.grad();
  }
}
/* forward kinematics of one tool from an explicit (state8, vel7 = v3 w3 gap_vel); with gnext8 != NULL also the adjoint
 * [g(state) 8 | g(vel) 7] */
template <class SimT>
static void probe_fk(SimT& S, int tool, const double* st, const double* vel, double* next8, const double* gnext8,
                     double* gout15) {
  typedef decltype(S.k.dt) TT;
  typedef Var<TT> VS;
  Tape<TT>& tp = tape<TT>();
  tp.clear();
  Pose<VS> P;
  for (int d = 0; d < 3; d++) P.pos[d] = VS::input((TT)st[d]);
  for (int d = 0; d < 4; d++) P.rot[d] = VS::input((TT)st[3 + d]);
  P.gap = VS::input((TT)st[7]);
  V3<VS> v{{VS::input((TT)vel[0]), VS::input((TT)vel[1]), VS::input((TT)vel[2])}};
  V3<VS> w{{VS::input((TT)vel[3]), VS::input((TT)vel[4]), VS::input((TT)vel[5])}};
  VS gv = VS::input((TT)vel[6]);
  Pose<VS> N = tool_fk<VS, TT>(S.tools[tool].c, P, v, w, gv);
  for (int d = 0; d < 3; d++) next8[d] = (double)val(N.pos[d]);
  for (int d = 0; d < 4; d++) next8[3 + d] = (double)val(N.rot[d]);
  next8[7] = (double)val(N.gap);
  if (!gnext8) return;
  tp.begin_reverse();
  for (int d = 0; d < 3; d++) N.pos[d].seed((TT)gnext8[d]);
  for (int d = 0; d < 4; d++) N.rot[d].seed((TT)gnext8[3 + d]);
  N.gap.seed((TT)gnext8[7]);
  tp.reverse();
  for (int d = 0; d < 3; d++) gout15[d] = (double)P.pos[d].grad();
  for (int d = 0; d < 4; d++) gout15[3 + d] = (double)P.rot[d].grad();
  gout15[7] = (double)P.gap.grad();
  for (int d = 0; d < 3; d++) gout15[8 + d] = (double)v[d].grad();
  for (int d = 0; d < 3; d++) gout15[11 + d] = (double)w[d].grad();
  gout15[14] = (double)gv.grad();
}
extern "C" {

void* orc_create(const orc_config* cfg, int use_f64) {
  Handle* h = new Handle{use_f64, nullptr, nullptr};
  if (use_f64)
    h->d = new Sim<double>(*cfg);
  else
    h->f = new Sim<float>(*cfg);
  return h;
}
void orc_destroy(void* h) {
  Handle* H = (Handle*)h;
  delete H->f;
  delete H->d;
  delete H;
}
void orc_set_scatter_noise(int seed, double ulps) {
  orc::g_scatter_noise_seed = seed;
  orc::g_scatter_noise_ulps = ulps;
}
void orc_set_pose_adjoint_atomics(int seed) { orc::g_pose_atomics_seed = seed; }
void orc_set_fast_math_noise(double amplitude, int salt) {
  ad::fast_math_noise() = amplitude;
  ad::fast_math_salt() = salt;
}
void orc_set_threads(int n) { omp_set_num_threads(n); }
void orc_set_svd_algorithm(int alg) { orc::g_svd_algorithm = alg; }
int orc_is_f64(void* h) { return ((Handle*)h)->f64; }
int orc_n_particles(void* h) {
  int r = 0;
  DISPATCH(h, r = S.n);
  return r;
}
void orc_initialize(void* h, int n_particles) { DISPATCH(h, S.n = n_particles); }
void orc_set_rand_num(void* h, const double* rn) {
  DISPATCH(h, if (S.npairs > 0) copy_in(S.rand_num, 0, rn, S.rand_num.size()));
}
void orc_set_material(void* h, const double* mu, const double* lam, const double* ys) {
  DISPATCH(h, {
    if (mu) copy_in(S.mu, 0, mu, S.n);
    if (lam) copy_in(S.lam, 0, lam, S.n);
    if (ys) copy_in(S.ys, 0, ys, S.n);
  });
}
void orc_set_tool_param(void* h, int tool, int which, double value) {
  DISPATCH(h, {
    auto& c = S.tools[tool].c;
    typedef decltype(c.friction) TT;
    if (which == 0) c.friction = (TT)value;
    else if (which == 1) c.softness = (TT)value;
    else if (which >= 2 && which <= 4) c.lo[which - 2] = (TT)value;
    else if (which >= 5 && which <= 7) c.hi[which - 5] = (TT)value;
  });
}
void orc_set_gravity(void* h, const double* g) {
  DISPATCH(h, for (int d = 0; d < 3; d++) S.k.gravity[d] = (decltype(S.k.dt))g[d]);
}
void orc_set_frame(void* h, int f, int n, const double* x, const double* v, const double* F, const double* C) {
  DISPATCH(h, {
    S.n = n;
    size_t o = (size_t)f * S.cap;
    copy_in(S.x, o * 3, x, (size_t)n * 3);
    copy_in(S.v, o * 3, v, (size_t)n * 3);
    copy_in(S.F, o * 9, F, (size_t)n * 9);
    copy_in(S.C, o * 9, C, (size_t)n * 9);
  });
}
void orc_get_frame(void* h, int f, double* x, double* v, double* F, double* C) {
  DISPATCH(h, {
    size_t o = (size_t)f * S.cap;
    if (x) copy_out(S.x, o * 3, x, (size_t)S.n * 3);
    if (v) copy_out(S.v, o * 3, v, (size_t)S.n * 3);
    if (F) copy_out(S.F, o * 9, F, (size_t)S.n * 9);
    if (C) copy_out(S.C, o * 9, C, (size_t)S.n * 9);
  });
}
void orc_set_tool_state(void* h, int f, int tool, const double* st) {
  DISPATCH(h, {
    auto& t = S.tools[tool];
    copy_in(t.pos, (size_t)f * 3, st, 3);
    copy_in(t.rot, (size_t)f * 4, st + 3, 4);
    copy_in(t.gap, (size_t)f, st + 7, 1);
  });
}
void orc_get_tool_state(void* h, int f, int tool, double* st) {
  DISPATCH(h, {
    auto& t = S.tools[tool];
    copy_out(t.pos, (size_t)f * 3, st, 3);
    copy_out(t.rot, (size_t)f * 4, st + 3, 4);
    copy_out(t.gap, (size_t)f, st + 7, 1);
  });
}
void orc_copyframe(void* h, int src, int dst) {
  DISPATCH(h, {
    size_t a = (size_t)src * S.cap, b = (size_t)dst * S.cap;
    std::copy(S.x.begin() + a * 3, S.x.begin() + (a + S.n) * 3, S.x.begin() + b * 3);
    std::copy(S.v.begin() + a * 3, S.v.begin() + (a + S.n) * 3, S.v.begin() + b * 3);
    std::copy(S.F.begin() + a * 9, S.F.begin() + (a + S.n) * 9, S.F.begin() + b * 9);
    std::copy(S.C.begin() + a * 9, S.C.begin() + (a + S.n) * 9, S.C.begin() + b * 9);
    for (auto& t : S.tools) {
      for (int d = 0; d < 3; d++) t.pos[dst * 3 + d] = t.pos[src * 3 + d];
      for (int d = 0; d < 4; d++) t.rot[dst * 4 + d] = t.rot[src * 4 + d];
      if (has_gap(t.c.type)) t.gap[dst] = t.gap[src];
    }
  });
}
void orc_set_action(void* h, int s, int nsub, const double* action) { DISPATCH(h, S.set_action(s, nsub, action)); }
void orc_substep(void* h, int f) { DISPATCH(h, S.substep(f)); }
void orc_substep_grad(void* h, int f) { DISPATCH(h, S.substep_grad(f)); }
void orc_set_velocity_grad(void* h, int s, int nsub) { DISPATCH(h, S.set_velocity_grad(s, nsub)); }
void orc_get_action_grad(void* h, int s, double* out) {
  DISPATCH(h, {
    int off = 0;
    for (auto& t : S.tools) {
      int ad = t.c.action_dim;
      if (ad > 0) copy_out(t.g_action, (size_t)s * ad, out + off, ad);
      off += ad;
    }
  });
}
void orc_zero_grad(void* h) { DISPATCH(h, S.zero_grad()); }
void orc_get_frame_grad(void* h, int f, double* gx, double* gv, double* gF, double* gC) {
  DISPATCH(h, {
    size_t o = (size_t)f * S.cap;
    if (gx) copy_out(S.gx, o * 3, gx, (size_t)S.n * 3);
    if (gv) copy_out(S.gv, o * 3, gv, (size_t)S.n * 3);
    if (gF) copy_out(S.gF, o * 9, gF, (size_t)S.n * 9);
    if (gC) copy_out(S.gC, o * 9, gC, (size_t)S.n * 9);
  });
}
void orc_add_frame_grad(void* h, int f, const double* gx, const double* gv, const double* gF, const double* gC) {
  DISPATCH(h, {
    size_t o = (size_t)f * S.cap;
    if (gx) add_in(S.gx, o * 3, gx, (size_t)S.n * 3);
    if (gv) add_in(S.gv, o * 3, gv, (size_t)S.n * 3);
    if (gF) add_in(S.gF, o * 9, gF, (size_t)S.n * 9);
    if (gC) add_in(S.gC, o * 9, gC, (size_t)S.n * 9);
  });
}
void orc_scale_frame_grad(void* h, int f, double alpha) {
  DISPATCH(h, {
    typedef decltype(S.k.dt) TT;
    size_t o = (size_t)f * S.cap;
    for (size_t i = 0; i < (size_t)S.n * 3; i++) {
      S.gx[o * 3 + i] *= (TT)alpha;
      S.gv[o * 3 + i] *= (TT)alpha;
    }
    for (size_t i = 0; i < (size_t)S.n * 9; i++) {
      S.gF[o * 9 + i] *= (TT)alpha;
      S.gC[o * 9 + i] *= (TT)alpha;
    }
    for (auto& t : S.tools) {
      for (int d = 0; d < 3; d++) t.g_pos[f * 3 + d] *= (TT)alpha;
      for (int d = 0; d < 4; d++) t.g_rot[f * 4 + d] *= (TT)alpha;
    }
  });
}
void orc_get_tool_grad(void* h, int f, int tool, double* g8) {
  DISPATCH(h, {
    auto& t = S.tools[tool];
    copy_out(t.g_pos, (size_t)f * 3, g8, 3);
    copy_out(t.g_rot, (size_t)f * 4, g8 + 3, 4);
    copy_out(t.g_gap, (size_t)f, g8 + 7, 1);
  });
}
void orc_add_tool_grad(void* h, int f, int tool, const double* g8) {
  DISPATCH(h, {
    auto& t = S.tools[tool];
    add_in(t.g_pos, (size_t)f * 3, g8, 3);
    add_in(t.g_rot, (size_t)f * 4, g8 + 3, 4);
    if (has_gap(t.c.type)) add_in(t.g_gap, (size_t)f, g8 + 7, 1);
  });
}
void orc_get_tool_vel_grad(void* h, int f, int tool, double* g7) {
  DISPATCH(h, {
    auto& t = S.tools[tool];
    copy_out(t.g_vel, (size_t)f * 3, g7, 3);
    copy_out(t.g_w, (size_t)f * 3, g7 + 3, 3);
    copy_out(t.g_gap_vel, (size_t)f, g7 + 6, 1);
  });
}
void orc_get_grid(void* h, double* v_in, double* v_out, double* m) {
  DISPATCH(h, {
    if (v_in) copy_out(S.grid_v_in, 0, v_in, (size_t)S.G * 3);
    if (v_out) copy_out(S.grid_v_out, 0, v_out, (size_t)S.G * 3);
    if (m) copy_out(S.grid_m, 0, m, (size_t)S.G);
  });
}
void orc_get_grid_grad(void* h, double* a, double* b, double* c) {
  DISPATCH(h, {
    if (a) copy_out(S.g_grid_v_in, 0, a, (size_t)S.G * 3);
    if (b) copy_out(S.g_grid_v_out, 0, b, (size_t)S.G * 3);
    if (c) copy_out(S.g_grid_m, 0, c, (size_t)S.G);
  });
}
void orc_get_svd(void* h, double* Ft, double* U, double* sig, double* V) {
  DISPATCH(h, {
    if (Ft) copy_out(S.Ftmp, 0, Ft, (size_t)S.n * 9);
    if (U) copy_out(S.U, 0, U, (size_t)S.n * 9);
    if (sig) copy_out(S.sig, 0, sig, (size_t)S.n * 3);
    if (V) copy_out(S.V, 0, V, (size_t)S.n * 9);
  });
}
void orc_cell_index(void* h, int f, int32_t* base, int32_t* key) {
  DISPATCH(h, {
    typedef decltype(S.k.dt) TT;
    for (int p = 0; p < S.n; p++) {
      int b[3];
      V3<TT> xx = S.template load_v3<TT>(S.x, (size_t)f * S.cap + p, false), fx, w[3];
      bspline(S.k, xx, b, fx, w);
      for (int d = 0; d < 3; d++) base[p * 3 + d] = b[d];
      if (key) key[p] = (b[0] * S.k.n + b[1]) * S.k.n + b[2];
    }
  });
}
void orc_occupancy(void* h, int f, uint8_t* occ) {
  DISPATCH(h, {
    typedef decltype(S.k.dt) TT;
    std::memset(occ, 0, (size_t)S.G);
    for (int p = 0; p < S.n; p++) {
      int b[3];
      V3<TT> xx = S.template load_v3<TT>(S.x, (size_t)f * S.cap + p, false), fx, w[3];
      bspline(S.k, xx, b, fx, w);
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          for (int l = 0; l < 3; l++) occ[((size_t)(b[0] + i) * S.k.n + (b[1] + j)) * S.k.n + (b[2] + l)] = 1;
    }
  });
}
void orc_get_collision_idx(void* h, int f, int32_t* idx) {
  DISPATCH(h, for (int c = 0; c < S.npairs; c++) idx[c] = S.collision_idx[(size_t)f * S.npairs + c]);
}
int orc_min_dist_cols(void* h) {
  int r = 0;
  DISPATCH(h, r = (int)S.dists.size());
  return r;
}
void orc_compute_min_dist(void* h, int f, double* out) {
  DISPATCH(h, {
    S.compute_min_dist(f);
    int nc = (int)S.dists.size();
    for (int p = 0; p < S.n; p++)
      for (int c = 0; c < nc; c++) out[(size_t)p * nc + c] = (double)S.dists[c][p];
  });
}
void orc_compute_min_dist_grad(void* h, int f, const double* gin) {
  DISPATCH(h, {
    typedef decltype(S.k.dt) TT;
    int nc = (int)S.dists.size();
    for (int p = 0; p < S.n; p++)
      for (int c = 0; c < nc; c++) S.g_dists[c][p] = (TT)gin[(size_t)p * nc + c];
    S.compute_min_dist_grad(f);
  });
}
void orc_compute_grid_m(void* h, int f, double* out) {
  DISPATCH(h, {
    S.compute_grid_m(f);
    copy_out(S.grid_m, 0, out, (size_t)S.G);
  });
}
void orc_compute_grid_m_grad(void* h, int f, const double* gin) {
  DISPATCH(h, {
    S.compute_grid_m(f);
    copy_in(S.g_grid_m, 0, gin, (size_t)S.G);
    S.compute_grid_m_grad(f);
  });
}
void orc_svd3(int use_f64, const double* F, double* U, double* sig, double* V) {
  if (use_f64) {
    if (orc::g_svd_algorithm == 1) svd3_eig_qr<double>(F, U, sig, V);
    else svd3<double>(F, U, sig, V);
  } else {
    float f[9], u[9], s[3], v[9];
    for (int i = 0; i < 9; i++) f[i] = (float)F[i];
    svd3<float>(f, u, s, v);
    for (int i = 0; i < 9; i++) {
      U[i] = u[i];
      V[i] = v[i];
    }
    for (int i = 0; i < 3; i++) sig[i] = s[i];
  }
}
double orc_tool_sdf(void* h, int tool, int f, const double* p) {
  double r = 0;
  DISPATCH(h, {
    typedef decltype(S.k.dt) TT;
    V3<TT> pp{{(TT)p[0], (TT)p[1], (TT)p[2]}};
    r = (double)tool_sdf<TT, TT>(S.tools[tool].c, S.template load_pose<TT>(tool, f, false), pp);
  });
  return r;
}
void orc_tool_normal(void* h, int tool, int f, const double* p, double* n) {
  DISPATCH(h, {
    typedef decltype(S.k.dt) TT;
    V3<TT> pp{{(TT)p[0], (TT)p[1], (TT)p[2]}};
    V3<TT> r = tool_normal<TT, TT>(S.tools[tool].c, S.template load_pose<TT>(tool, f, false), pp);
    for (int d = 0; d < 3; d++) n[d] = (double)r[d];
  });
}
void orc_tool_collide(void* h, int tool, int f, const double* p, const double* v_in, double* v_out) {
  DISPATCH(h, {
    typedef decltype(S.k.dt) TT;
    V3<TT> pp{{(TT)p[0], (TT)p[1], (TT)p[2]}}, vv{{(TT)v_in[0], (TT)v_in[1], (TT)v_in[2]}};
    V3<TT> r = tool_collide<TT, TT>(S.tools[tool].c, S.template load_pose<TT>(tool, f, false),
                                    S.template load_pose<TT>(tool, f + 1, false), pp, vv, S.k.dt);
    for (int d = 0; d < 3; d++) v_out[d] = (double)r[d];
  });
}
void orc_tool_probe_grad(void* h, int tool, int f, int what, const double* p, const double* v_in, const double* gout,
                         double* out22) {
  DISPATCH(h, { probe_grad(S, tool, f, what, p, v_in, gout, out22); });
}
void orc_tool_probe_fk(void* h, int tool, const double* state8, const double* vel7, double* next8,
                       const double* gnext8, double* gout15) {
  DISPATCH(h, { probe_fk(S, tool, state8, vel7, next8, gnext8, gout15); });
}
}
