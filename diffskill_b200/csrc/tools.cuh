// Rigid-tool signed distance fields, contact response and their hand-derived adjoints.
//
// Replaces the Taichi ti.funcs of plb/engine/primitive/primive_base.py:75-120
// (sdf / normal / collider_v / collide) and primitives.py (Capsule :54-66, Box :374-393,
// Gripper :471-536, Prism :711-731, Knife :769-807), plus the `.grad` code Taichi's
// autodiff derives from them (call site mpm_simulator.py:336).
//
// Forward values use non-contracting `_rn` arithmetic in the reference's operation
// order: the finite-difference normals amplify 1-ulp differences by 0.5/1e-4.
// Adjoints follow Taichi's per-op rules: min(a,b) -> a iff a<b; max(a,b) -> a iff b<a;
// abs -> sgn; predicates and selects carry no gradient; the FD stencil is
// differentiated through (six SDF-gradient evaluations).
#pragma once
#include "../../include/diffskill_mpm.h"
#include "mpm_math.cuh"

struct ToolParams {
  int type, action_dim;
  float action_scale[8];
  float friction, softness;
  float lo[3], hi[3];
  float size[3];
  float h, half_h, r;
  float prism_h0, prism_h1;
  Q4 prot_inv;  // normalised conjugate of prot (computed on device in fp32 like the reference)
  Q4 prot;
  float min_gap, max_gap;
  float bound_r;  // radius of a sphere around the frame origin that contains the shape (tile culling)
};

struct Pose {  // position[f], rotation[f], gap[f]
  float3 p;
  Q4 q;
  float gap;
};
struct PoseAdj {
  float3 p;
  Q4 q;
  float gap;
};
DSK_DEV PoseAdj pose_adj_zero() {
  PoseAdj a;
  a.p = f3(0, 0, 0);
  a.q.w = a.q.x = a.q.y = a.q.z = 0.f;
  a.gap = 0.f;
  return a;
}
DSK_DEV Pose load_pose(const float* s) {
  Pose P;
  P.p = f3(s[0], s[1], s[2]);
  P.q.w = s[3];
  P.q.x = s[4];
  P.q.y = s[5];
  P.q.z = s[6];
  P.gap = s[7];
  return P;
}

// a rigid frame: origin + (unnormalised) rotation; tools use (position, rotation), gripper
// jaws use (get_pos(flag), rotation)
struct Frame {
  float3 o;
  Q4 q;
  float aux;   // shape parameter that lives in the pose: gap[f] of a Chopsticks tool frame (0 otherwise)
};
struct FrameAdj {
  float3 o;
  Q4 q;
  float aux;
};
DSK_DEV FrameAdj frame_adj_zero() {
  FrameAdj a;
  a.o = f3(0, 0, 0);
  a.q.w = a.q.x = a.q.y = a.q.z = 0.f;
  a.aux = 0.f;
  return a;
}

// SDF_SPHERE (primitives.py:23-41): the reference evaluates length(p - position) - radius in WORLD space; here it is
// |R^-1 (p - position)| - radius in the tool frame like every other shape -- the same number up to an fp32 rounding of
// the rotation, and the rotation adjoint it produces is zero to the same rounding.
// SDF_CYLINDER (primitives.py:302-336): T.h = radial, T.r = axial half extent.  SDF_TORUS (primitives.py:337-365):
// major radius tx in T.h, minor radius ty in T.r.  Both have analytic normals.
// SDF_CHOPSTICKS (primitives.py:245-261): two capsules at -+gap/2 along local x, shifted by h/2 along local y, inside the ONE
// tool frame; sdf = min of the two, normal = the nearer one's.  The gap travels in Frame::aux; the local_* functions are never
// called with this kind -- the frame-level functions pick the stick and call them with SDF_CAPSULE on the shifted point.
enum { SDF_CAPSULE = 0, SDF_BOX = 1, SDF_KNIFE = 2, SDF_SPHERE = 3, SDF_CYLINDER = 4, SDF_TORUS = 5, SDF_CHOPSTICKS = 6 };
// the local shape of a tool's contact frame(s): the tool itself, or one jaw of a Gripper (box jaws, primitives.py:475-483)
// / Gripper2 (capsule jaws, primitives.py:607-615)
DSK_DEV int sdf_kind(int tool_type) {
  switch (tool_type) {
    case DSK_TOOL_CAPSULE:
    case DSK_TOOL_ROLLINGPIN_EXT:
    case DSK_TOOL_ROLLINGPIN:
    case DSK_TOOL_GRIPPER2: return SDF_CAPSULE;
    case DSK_TOOL_KNIFE: return SDF_KNIFE;
    case DSK_TOOL_SPHERE: return SDF_SPHERE;
    case DSK_TOOL_CYLINDER: return SDF_CYLINDER;
    case DSK_TOOL_TORUS: return SDF_TORUS;
    case DSK_TOOL_CHOPSTICKS: return SDF_CHOPSTICKS;
    default: return SDF_BOX;   // Box, Gripper
  }
}
// tools with two jaws applied sequentially, a gap state and a 7-D action (primitives.py:428, :576)
DSK_DEV bool is_gripper(int tool_type) { return tool_type == DSK_TOOL_GRIPPER || tool_type == DSK_TOOL_GRIPPER2; }
// tools with a gap state, an 8-float state and a 7-D action: the grippers and Chopsticks (primitives.py:218)
DSK_DEV bool has_gap(int tool_type) { return is_gripper(tool_type) || tool_type == DSK_TOOL_CHOPSTICKS; }
DSK_DEV bool is_rollingpin(int tool_type) { return tool_type == DSK_TOOL_ROLLINGPIN_EXT || tool_type == DSK_TOOL_ROLLINGPIN; }

// ---- local SDFs (value) ----------------------------------------------------------------------
DSK_DEV float len14_rn(float3 v) { return __fsqrt_rn(add_rn(dot_rn(v, v), 1e-14f)); }
DSK_DEV float len8_rn(float3 v) { return __fsqrt_rn(add_rn(dot_rn(v, v), 1e-8f)); }

DSK_DEV float box_sdf(const ToolParams& T, float3 p) {  // primitives.py:374-380
  float q0 = sub_rn(fabsf(p.x), T.size[0]), q1 = sub_rn(fabsf(p.y), T.size[1]), q2 = sub_rn(fabsf(p.z), T.size[2]);
  float3 mq = f3(tmax(q0, 0.f), tmax(q1, 0.f), tmax(q2, 0.f));
  float out = len14_rn(mq);
  return add_rn(out, tmin(tmax(q0, tmax(q1, q2)), 0.f));
}
DSK_DEV float3 capsule_p2(const ToolParams& T, float3 p) {  // primitives.py:56-58
  float y = add_rn(p.y, T.half_h);
  y = sub_rn(y, tmin(tmax(y, 0.f), T.h));
  return f3(p.x, y, p.z);
}
DSK_DEV float prism_sdf(const ToolParams& T, float3 p0) {  // primitives.py:711-718
  float3 p = qrot_rn(T.prot_inv, p0);
  float a = add_rn(mul_rn(fabsf(p.x), 0.866025f), mul_rn(p.y, 0.5f));
  return tmax(sub_rn(fabsf(p.z), T.prism_h1), sub_rn(tmax(a, -p.y), mul_rn(T.prism_h0, 0.5f)));
}
DSK_DEV float len14_2_rn(float a, float b) { return __fsqrt_rn(add_rn(add_rn(mul_rn(a, a), mul_rn(b, b)), 1e-14f)); }
DSK_DEV float cylinder_sdf(const ToolParams& T, float3 p) {  // primitives.py:309-313
  float l = len14_2_rn(p.x, p.z);
  float d0 = sub_rn(fabsf(l), T.h), d1 = sub_rn(fabsf(p.y), T.r);
  return add_rn(tmin(tmax(d0, d1), 0.f), len14_2_rn(tmax(d0, 0.f), tmax(d1, 0.f)));
}
DSK_DEV float torus_sdf(const ToolParams& T, float3 p) {  // primitives.py:344-347
  float q0 = sub_rn(len14_2_rn(p.x, p.z), T.h);
  return sub_rn(len14_2_rn(q0, p.y), T.r);
}
DSK_DEV float local_sdf(const ToolParams& T, int kind, float3 p) {
  if (kind == SDF_CAPSULE) return sub_rn(len14_rn(capsule_p2(T, p)), T.r);
  if (kind == SDF_BOX) return box_sdf(T, p);
  if (kind == SDF_SPHERE) return sub_rn(len14_rn(p), T.r);   // primitives.py:28-30
  if (kind == SDF_CYLINDER) return cylinder_sdf(T, p);
  if (kind == SDF_TORUS) return torus_sdf(T, p);
  return tmax(prism_sdf(T, p), box_sdf(T, p));  // primitives.py:769-773
}

// ---- local SDFs (gradient wrt the local point, Taichi AD rules) ---------------------------------
DSK_DEV float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
DSK_DEV float3 box_sdf_grad(const ToolParams& T, float3 p) {
  float q0 = fabsf(p.x) - T.size[0], q1 = fabsf(p.y) - T.size[1], q2 = fabsf(p.z) - T.size[2];
  float3 mq = f3(tmax(q0, 0.f), tmax(q1, 0.f), tmax(q2, 0.f));
  float L = sqrtf(dot(mq, mq) + 1e-14f);
  // d len / d q_i = mq_i / L, only where 0 < q_i
  float g0 = (0.f < q0) ? mq.x / L : 0.f;
  float g1 = (0.f < q1) ? mq.y / L : 0.f;
  float g2 = (0.f < q2) ? mq.z / L : 0.f;
  // min(max(q0, max(q1,q2)), 0)
  float m12 = tmax(q1, q2);
  float m012 = tmax(q0, m12);
  if (m012 < 0.f) {
    if (m12 < q0) g0 += 1.f;
    else if (q2 < q1) g1 += 1.f;
    else g2 += 1.f;
  }
  return f3(g0 * sgnf(p.x), g1 * sgnf(p.y), g2 * sgnf(p.z));
}
DSK_DEV float3 capsule_sdf_grad(const ToolParams& T, float3 p) {
  float y = p.y + T.half_h;
  float t = tmax(y, 0.f);
  float dyy = 1.f - (((0.f < y) && (t < T.h)) ? 1.f : 0.f);
  float3 p2 = capsule_p2(T, p);
  float L = sqrtf(dot(p2, p2) + 1e-14f);
  return f3(p2.x / L, dyy * p2.y / L, p2.z / L);
}
DSK_DEV float3 prism_sdf_grad(const ToolParams& T, float3 p0) {
  float3 p = qrot_rn(T.prot_inv, p0);
  float a = fabsf(p.x) * 0.866025f + p.y * 0.5f;
  float b = -p.y;
  float inner = tmax(a, b) - T.prism_h0 * 0.5f;
  float outer_l = fabsf(p.z) - T.prism_h1;
  float3 g = f3(0, 0, 0);
  if (inner < outer_l) {
    g.z = sgnf(p.z);
  } else if (b < a) {
    g.x = 0.866025f * sgnf(p.x);
    g.y = 0.5f;
  } else {
    g.y = -1.f;
  }
  Q4 gq = {0, 0, 0, 0};
  float3 gp0 = f3(0, 0, 0);
  qrot_adj(T.prot_inv, p0, g, gq, gp0);
  return gp0;
}
DSK_DEV float3 cylinder_sdf_grad(const ToolParams& T, float3 p) {
  float l = sqrtf(p.x * p.x + p.z * p.z + 1e-14f);
  float d0 = l - T.h, d1 = fabsf(p.y) - T.r;   // |l| = l: l > 0
  float g0 = 0.f, g1 = 0.f;
  if (tmax(d0, d1) < 0.f) {  // min(max(d0,d1), 0): the max gets it iff it is < 0; max(a,b) -> a iff b < a
    if (d1 < d0) g0 += 1.f;
    else g1 += 1.f;
  }
  float m0 = tmax(d0, 0.f), m1 = tmax(d1, 0.f);
  float L = sqrtf(m0 * m0 + m1 * m1 + 1e-14f);
  if (0.f < d0) g0 += m0 / L;
  if (0.f < d1) g1 += m1 / L;
  return f3(g0 * p.x / l, g1 * sgnf(p.y), g0 * p.z / l);
}
DSK_DEV float3 torus_sdf_grad(const ToolParams& T, float3 p) {
  float l = sqrtf(p.x * p.x + p.z * p.z + 1e-14f);
  float q0 = l - T.h;
  float lq = sqrtf(q0 * q0 + p.y * p.y + 1e-14f);
  float gl = q0 / lq;
  return f3(gl * p.x / l, p.y / lq, gl * p.z / l);
}
DSK_DEV float3 local_sdf_grad(const ToolParams& T, int kind, float3 p) {
  if (kind == SDF_CAPSULE) return capsule_sdf_grad(T, p);
  if (kind == SDF_BOX) return box_sdf_grad(T, p);
  if (kind == SDF_SPHERE) return (1.f / sqrtf(dot(p, p) + 1e-14f)) * p;
  if (kind == SDF_CYLINDER) return cylinder_sdf_grad(T, p);
  if (kind == SDF_TORUS) return torus_sdf_grad(T, p);
  float a = prism_sdf(T, p), b = box_sdf(T, p);
  return (b < a) ? prism_sdf_grad(T, p) : box_sdf_grad(T, p);
}

// ---- local normals ---------------------------------------------------------------------------------
#define DSK_FD_D 1e-4f
// Cylinder._normal before the final normalize (primitives.py:315-328)
DSK_DEV float3 cylinder_n3(const ToolParams& T, float3 p) {
  float l = len14_2_rn(p.x, p.z);
  float d0 = sub_rn(l, T.h), d1 = sub_rn(fabsf(p.y), T.r);
  float f = d0 > d1 ? 1.f : 0.f;
  float inside = tmax(d0, d1) <= 0.f ? 1.f : 0.f;
  float n20 = add_rn(tmax(d0, 0.f), mul_rn(inside, f)), n21 = add_rn(tmax(d1, 0.f), mul_rn(inside, sub_rn(1.f, f)));
  float ln = len14_2_rn(n20, n21);
  float a = __fdiv_rn(n20, ln), b = __fdiv_rn(n21, ln);
  float sy = p.y >= 0.f ? 1.f : -1.f;
  return f3(mul_rn(__fdiv_rn(p.x, l), a), mul_rn(b, sy), mul_rn(__fdiv_rn(p.z, l), a));
}
// adjoint of n3 = cylinder_n3(p) wrt p; the casts (f, inside, sy) carry no gradient
DSK_DEV float3 cylinder_n3_adj(const ToolParams& T, float3 p, float3 g) {
  float l = sqrtf(p.x * p.x + p.z * p.z + 1e-14f);
  float d0 = l - T.h, d1 = fabsf(p.y) - T.r;
  float f = d0 > d1 ? 1.f : 0.f;
  float inside = tmax(d0, d1) <= 0.f ? 1.f : 0.f;
  float n20 = tmax(d0, 0.f) + inside * f, n21 = tmax(d1, 0.f) + inside * (1.f - f);
  float ln = sqrtf(n20 * n20 + n21 * n21 + 1e-14f);
  float a = n20 / ln;
  float sy = p.y >= 0.f ? 1.f : -1.f;
  float p20 = p.x / l, p21 = p.z / l;
  float gp20 = g.x * a, gp21 = g.z * a;
  float ga = g.x * p20 + g.z * p21, gb = g.y * sy;
  float gln = -(ga * n20 + gb * n21) / (ln * ln);
  float gn20 = ga / ln + gln * n20 / ln, gn21 = gb / ln + gln * n21 / ln;
  float gl = (0.f < d0) ? gn20 : 0.f;          // max(d0, 0): d0 gets it iff 0 < d0
  float gy = ((0.f < d1) ? gn21 : 0.f) * sgnf(p.y);
  gl -= (gp20 * p.x + gp21 * p.z) / (l * l);   // p2 = (x, z) / l
  return f3(gp20 / l + gl * p.x / l, gy, gp21 / l + gl * p.z / l);
}
// Torus._normal before the final normalize (primitives.py:349-357)
DSK_DEV float3 torus_n3(const ToolParams& T, float3 p) {
  float l = len14_2_rn(p.x, p.z);
  float q0 = sub_rn(l, T.h);
  float lq = len14_2_rn(q0, p.y);
  float n20 = __fdiv_rn(q0, lq), n21 = __fdiv_rn(p.y, lq);
  return f3(mul_rn(__fdiv_rn(p.x, l), n20), n21, mul_rn(__fdiv_rn(p.z, l), n20));
}
DSK_DEV float3 torus_n3_adj(const ToolParams& T, float3 p, float3 g) {
  float l = sqrtf(p.x * p.x + p.z * p.z + 1e-14f);
  float q0 = l - T.h, q1 = p.y;
  float lq = sqrtf(q0 * q0 + q1 * q1 + 1e-14f);
  float n20 = q0 / lq;
  float x20 = p.x / l, x21 = p.z / l;
  float gx20 = g.x * n20, gx21 = g.z * n20;
  float gn20 = g.x * x20 + g.z * x21, gn21 = g.y;
  float glq = -(gn20 * q0 + gn21 * q1) / (lq * lq);
  float gq0 = gn20 / lq + glq * q0 / lq, gq1 = gn21 / lq + glq * q1 / lq;
  float gl = gq0 - (gx20 * p.x + gx21 * p.z) / (l * l);
  return f3(gx20 / l + gl * p.x / l, gq1, gx21 / l + gl * p.z / l);
}
DSK_DEV float3 local_normal_raw(const ToolParams& T, int kind, float3 p, float& L) {
  // returns the un-normalised n and its length (eps 1e-14)
  float3 n;
  if (kind == SDF_CAPSULE) {
    n = capsule_p2(T, p);
  } else if (kind == SDF_SPHERE) {  // primitives.py:32-34: normalize(p)
    n = p;
  } else if (kind == SDF_CYLINDER) {
    n = cylinder_n3(T, p);
  } else if (kind == SDF_TORUS) {
    n = torus_n3(T, p);
  } else {  // primitives.py:382-393 / 796-807
    const float d = DSK_FD_D;
    const float c = __fdiv_rn(0.5f, d);
    n.x = mul_rn(c, sub_rn(local_sdf(T, kind, f3(add_rn(p.x, d), p.y, p.z)), local_sdf(T, kind, f3(sub_rn(p.x, d), p.y, p.z))));
    n.y = mul_rn(c, sub_rn(local_sdf(T, kind, f3(p.x, add_rn(p.y, d), p.z)), local_sdf(T, kind, f3(p.x, sub_rn(p.y, d), p.z))));
    n.z = mul_rn(c, sub_rn(local_sdf(T, kind, f3(p.x, p.y, add_rn(p.z, d))), local_sdf(T, kind, f3(p.x, p.y, sub_rn(p.z, d)))));
  }
  L = len14_rn(n);
  return n;
}
DSK_DEV float3 local_normal(const ToolParams& T, int kind, float3 p) {
  float L;
  float3 n = local_normal_raw(T, kind, p, L);
  return f3(__fdiv_rn(n.x, L), __fdiv_rn(n.y, L), __fdiv_rn(n.z, L));
}
// adjoint of N = local_normal(p) wrt p
DSK_DEV float3 local_normal_adj(const ToolParams& T, int kind, float3 p, float3 gN) {
  float L;
  float3 n = local_normal_raw(T, kind, p, L);
  // N = n / L, L = sqrt(n.n + eps)
  float3 gn = (1.f / L) * gN - (dot(n, gN) / (L * L * L)) * n;
  if (kind == SDF_CAPSULE) {
    float y = p.y + T.half_h;
    float t = tmax(y, 0.f);
    float dyy = 1.f - (((0.f < y) && (t < T.h)) ? 1.f : 0.f);
    return f3(gn.x, dyy * gn.y, gn.z);
  }
  if (kind == SDF_SPHERE) return gn;   // n = p
  if (kind == SDF_CYLINDER) return cylinder_n3_adj(T, p, gn);
  if (kind == SDF_TORUS) return torus_n3_adj(T, p, gn);
  const float d = DSK_FD_D;
  const float c = 0.5f / d;
  float3 gp = f3(0, 0, 0);
  gp += (c * gn.x) * (local_sdf_grad(T, kind, f3(p.x + d, p.y, p.z)) - local_sdf_grad(T, kind, f3(p.x - d, p.y, p.z)));
  gp += (c * gn.y) * (local_sdf_grad(T, kind, f3(p.x, p.y + d, p.z)) - local_sdf_grad(T, kind, f3(p.x, p.y - d, p.z)));
  gp += (c * gn.z) * (local_sdf_grad(T, kind, f3(p.x, p.y, p.z + d)) - local_sdf_grad(T, kind, f3(p.x, p.y, p.z - d)));
  return gp;
}

// the same with the un-normalised normal n and its length L already known (cached by the forward geometry pass)
DSK_DEV float3 local_normal_adj_cached(const ToolParams& T, int kind, float3 p, float3 n, float L, float3 gN) {
  float3 gn = (1.f / L) * gN - (dot(n, gN) / (L * L * L)) * n;
  if (kind == SDF_CAPSULE) {
    float y = p.y + T.half_h;
    float t = tmax(y, 0.f);
    float dyy = 1.f - (((0.f < y) && (t < T.h)) ? 1.f : 0.f);
    return f3(gn.x, dyy * gn.y, gn.z);
  }
  if (kind == SDF_SPHERE) return gn;   // n = p
  if (kind == SDF_CYLINDER) return cylinder_n3_adj(T, p, gn);
  if (kind == SDF_TORUS) return torus_n3_adj(T, p, gn);
  const float d = DSK_FD_D;
  const float c = 0.5f / d;
  float3 gp = f3(0, 0, 0);
  gp += (c * gn.x) * (local_sdf_grad(T, kind, f3(p.x + d, p.y, p.z)) - local_sdf_grad(T, kind, f3(p.x - d, p.y, p.z)));
  gp += (c * gn.y) * (local_sdf_grad(T, kind, f3(p.x, p.y + d, p.z)) - local_sdf_grad(T, kind, f3(p.x, p.y - d, p.z)));
  gp += (c * gn.z) * (local_sdf_grad(T, kind, f3(p.x, p.y, p.z + d)) - local_sdf_grad(T, kind, f3(p.x, p.y, p.z - d)));
  return gp;
}

// ---- world <-> tool frame -----------------------------------------------------------------------
DSK_DEV float3 inv_trans(const Frame& F, float3 p) {  // utils.py:50-54
  return qrot_rn(qconj_normalized_rn(F.q), sub3_rn(p, F.o));
}
// adjoint of pl = inv_trans(F, p) given g(pl); accumulates into gF and gp
DSK_DEV void inv_trans_adj(const Frame& F, float3 p, float3 gpl, FrameAdj& gF, float3& gp) {
  Q4 qn = qconj_normalized_rn(F.q);
  float3 d = p - F.o;
  Q4 gqn = {0, 0, 0, 0};
  float3 gd = f3(0, 0, 0);
  qrot_adj(qn, d, gpl, gqn, gd);
  qconj_normalized_adj(F.q, gqn, gF.q);
  gF.o -= gd;
  gp += gd;
}

// Chopsticks: the two candidate points in capsule coordinates and their signed distances (primitives.py:245-250)
struct ChopEval {
  float3 pa, pb;
  float a, b;
};
DSK_DEV ChopEval chop_eval(const ToolParams& T, float aux, float3 pl) {
  ChopEval c;
  float y = sub_rn(pl.y, -T.half_h);            // grid_pos - (0, -h/2, 0)
  float half = __fdiv_rn(aux, 2.f);
  c.pa = f3(sub_rn(pl.x, half), y, pl.z);       // p - delta
  c.pb = f3(add_rn(pl.x, half), y, pl.z);       // p + delta
  c.a = local_sdf(T, SDF_CAPSULE, c.pa);
  c.b = local_sdf(T, SDF_CAPSULE, c.pb);
  return c;
}
DSK_DEV float frame_sdf(const ToolParams& T, int kind, const Frame& F, float3 p) {
  if (kind == SDF_CHOPSTICKS) {
    ChopEval c = chop_eval(T, F.aux, inv_trans(F, p));
    return tmin(c.a, c.b);
  }
  return local_sdf(T, kind, inv_trans(F, p));
}
DSK_DEV void frame_sdf_adj(const ToolParams& T, int kind, const Frame& F, float3 p, float gd, FrameAdj& gF,
                           float3& gp) {
  float3 pl = inv_trans(F, p);
  if (kind == SDF_CHOPSTICKS) {   // min(a, b): a gets the adjoint iff a < b
    ChopEval c = chop_eval(T, F.aux, pl);
    bool first = c.a < c.b;
    float3 g = gd * local_sdf_grad(T, SDF_CAPSULE, first ? c.pa : c.pb);
    gF.aux += (first ? -0.5f : 0.5f) * g.x;
    inv_trans_adj(F, p, g, gF, gp);
    return;
  }
  float3 g = gd * local_sdf_grad(T, kind, pl);
  inv_trans_adj(F, p, g, gF, gp);
}
DSK_DEV float3 frame_normal(const ToolParams& T, int kind, const Frame& F, float3 p) {  // primive_base.py:80-85
  if (kind == SDF_CHOPSTICKS) {   // primitives.py:253-261: the nearer stick's normal (a <= b)
    ChopEval c = chop_eval(T, F.aux, inv_trans(F, p));
    return qrot_rn(F.q, local_normal(T, SDF_CAPSULE, c.a <= c.b ? c.pa : c.pb));
  }
  return qrot_rn(F.q, local_normal(T, kind, inv_trans(F, p)));
}
DSK_DEV void frame_normal_adj(const ToolParams& T, int kind, const Frame& F, float3 p, float3 gD, FrameAdj& gF,
                              float3& gp) {
  float3 pl = inv_trans(F, p);
  if (kind == SDF_CHOPSTICKS) {
    ChopEval c = chop_eval(T, F.aux, pl);
    bool first = c.a <= c.b;
    float3 ps = first ? c.pa : c.pb;
    float3 Nl = local_normal(T, SDF_CAPSULE, ps);
    float3 gNl = f3(0, 0, 0);
    qrot_adj(F.q, Nl, gD, gF.q, gNl);
    float3 gpl = local_normal_adj(T, SDF_CAPSULE, ps, gNl);
    gF.aux += (first ? -0.5f : 0.5f) * gpl.x;
    inv_trans_adj(F, p, gpl, gF, gp);
    return;
  }
  float3 Nl = local_normal(T, kind, pl);
  float3 gNl = f3(0, 0, 0);
  qrot_adj(F.q, Nl, gD, gF.q, gNl);
  float3 gpl = local_normal_adj(T, kind, pl, gNl);
  inv_trans_adj(F, p, gpl, gF, gp);
}
// adjoint of the local part of a contact's geometry -- normal N = n / L at the point the forward pass took it and
// dist = sdf -- w.r.t. the tool-local node position pl (and, for Chopsticks, the gap in Frame::aux); n, L cached by
// contact_geometry.  Shared by the grid_op.grad kernels and the CPU twin.
DSK_DEV float3 contact_local_adj(const ToolParams& T, int kind, float aux, float3 pl, float3 nraw, float L, float3 gNl,
                                 float gdist, float& gaux) {
  if (kind == SDF_CHOPSTICKS) {
    ChopEval e = chop_eval(T, aux, pl);
    bool fn = e.a <= e.b, fd = e.a < e.b;   // normal: the select of primitives.py:259; sdf: min(a, b) -> a iff a < b
    float3 g1 = local_normal_adj_cached(T, SDF_CAPSULE, fn ? e.pa : e.pb, nraw, L, gNl);
    float3 g2 = gdist * local_sdf_grad(T, SDF_CAPSULE, fd ? e.pa : e.pb);
    gaux += (fn ? -0.5f : 0.5f) * g1.x + (fd ? -0.5f : 0.5f) * g2.x;
    return g1 + g2;
  }
  float3 gpl = local_normal_adj_cached(T, kind, pl, nraw, L, gNl);
  gpl += gdist * local_sdf_grad(T, kind, pl);
  return gpl;
}
// collider_v, primive_base.py:87-94
DSK_DEV float3 frame_collider_v(const Frame& F0, const Frame& F1, float3 p, float dt) {
  float3 rel = qrot_rn(qconj_normalized_rn(F0.q), sub3_rn(p, F0.o));
  float3 np = add3_rn(qrot_rn(F1.q, rel), F1.o);
  return f3(__fdiv_rn(sub_rn(np.x, p.x), dt), __fdiv_rn(sub_rn(np.y, p.y), dt), __fdiv_rn(sub_rn(np.z, p.z), dt));
}
DSK_DEV void frame_collider_v_adj(const Frame& F0, const Frame& F1, float3 p, float dt, float3 gcv, FrameAdj& g0,
                                  FrameAdj& g1) {
  float3 rel = qrot_rn(qconj_normalized_rn(F0.q), sub3_rn(p, F0.o));
  float3 gnp = (1.f / dt) * gcv;
  g1.o += gnp;
  float3 grel = f3(0, 0, 0);
  qrot_adj(F1.q, rel, gnp, g1.q, grel);
  float3 gp_unused = f3(0, 0, 0);
  inv_trans_adj(F0, p, grel, g0, gp_unused);
}

// ---- contact response ------------------------------------------------------------------------------
// Primitive.collide (primive_base.py:96-120, eps14 = false) / Gripper.collide2 (primitives.py:513-536, eps14 = true)
DSK_DEV bool contact_active(float dist, float softness, float& influence) {
  influence = tmin(expf(-dist * softness), 1.f);
  return (softness > 0.f && influence > 0.1f) || dist <= 0.f;
}
DSK_DEV float3 contact_response(float3 v, float3 D, float3 cv, float influence, float friction, bool eps14) {
  float3 u = v - cv;
  float nc = dot(u, D);
  float mn = tmin(nc, 0.f);
  float3 t = u - mn * D;
  float tn = sqrtf(dot(t, t) + (eps14 ? 1e-14f : 1e-8f));
  float mx = tmax(0.f, tn + nc * friction);
  bool flag = (nc < 0.f) && (sqrtf(dot(t, t)) > 1e-30f);
  float3 t2 = flag ? f3(t.x / tn * mx, t.y / tn * mx, t.z / tn * mx) : t;
  return cv + u * (1.f - influence) + t2 * influence;
}
// forward of one contact against frame (F0 at f, F1 at f+1)
DSK_DEV float3 contact_forward(const ToolParams& T, int kind, const Frame& F0, const Frame& F1, float3 p, float3 v,
                               float dt, bool eps14) {
  float dist = frame_sdf(T, kind, F0, p);
  float influence;
  if (contact_active(dist, T.softness, influence)) {
    float3 D = frame_normal(T, kind, F0, p);
    float3 cv = frame_collider_v(F0, F1, p, dt);
    v = contact_response(v, D, cv, influence, T.friction, eps14);
  }
  return v;
}
// adjoint: given g(v_out) returns g(v_in) and accumulates frame adjoints
DSK_DEV float3 contact_adjoint(const ToolParams& T, int kind, const Frame& F0, const Frame& F1, float3 p, float3 v,
                               float dt, bool eps14, float3 gout, FrameAdj& g0, FrameAdj& g1) {
  float dist = frame_sdf(T, kind, F0, p);
  float influence;
  if (!contact_active(dist, T.softness, influence)) return gout;
  float3 D = frame_normal(T, kind, F0, p);
  float3 cv = frame_collider_v(F0, F1, p, dt);
  float friction = T.friction;
  // recompute forward intermediates
  float3 u = v - cv;
  float nc = dot(u, D);
  float mn = tmin(nc, 0.f);
  float3 t = u - mn * D;
  float tn = sqrtf(dot(t, t) + (eps14 ? 1e-14f : 1e-8f));
  float a2 = tn + nc * friction;
  float mx = tmax(0.f, a2);
  bool flag = (nc < 0.f) && (sqrtf(dot(t, t)) > 1e-30f);
  float3 q = (1.f / tn) * t;
  float3 t2 = flag ? mx * q : t;
  // out = cv + u*(1-infl) + t2*infl
  float3 gcv = gout;
  float3 gu = (1.f - influence) * gout;
  float ginfl = dot(gout, t2 - u);
  float3 gt2 = influence * gout;
  float3 gt = f3(0, 0, 0);
  float gnc = 0.f;
  if (flag) {
    // t2 = (t/tn) * mx
    float3 gq = mx * gt2;
    float gmx = dot(gt2, q);
    float ga2 = (a2 < 0.f) ? 0.f : gmx;  // max(0, a2): rhs gets it unless a2 < 0
    float gtn = ga2 - dot(gq, t) / (tn * tn);
    gt += (1.f / tn) * gq;
    gnc += ga2 * friction;
    gt += (gtn / tn) * t;  // tn = sqrt(t.t + eps)
  } else {
    gt += gt2;
  }
  // t = u - mn*D
  gu += gt;
  float gmn = -dot(gt, D);
  float3 gD = (-mn) * gt;
  if (nc < 0.f) gnc += gmn;  // min(nc, 0)
  // nc = u.D
  gu += gnc * D;
  gD += gnc * u;
  // u = v - cv
  float3 gv = gu;
  gcv -= gu;
  // influence = min(exp(-dist*softness), 1)
  float e = expf(-dist * T.softness);
  float gdist = (e < 1.f) ? (-T.softness * e * ginfl) : 0.f;
  float3 gp_unused = f3(0, 0, 0);
  frame_normal_adj(T, kind, F0, p, gD, g0, gp_unused);
  frame_collider_v_adj(F0, F1, p, dt, gcv, g0, g1);
  frame_sdf_adj(T, kind, F0, p, gdist, g0, gp_unused);
  return gv;
}

// ---- tool level: frames of a tool, gripper jaws ----------------------------------------------------
DSK_DEV Frame tool_frame(const Pose& P) {
  Frame F;
  F.o = P.p;
  F.q = P.q;
  F.aux = P.gap;   // read by the Chopsticks shape only
  return F;
}
DSK_DEV Frame jaw_frame(const Pose& P, float flag) {  // Gripper.get_pos, primitives.py:471-473
  Frame F;
  float off = mul_rn(__fdiv_rn(P.gap, 2.f), flag);
  F.o = add3_rn(P.p, qrot_rn(P.q, f3(off, 0.f, 0.f)));
  F.q = P.q;
  F.aux = 0.f;
  return F;
}
DSK_DEV void jaw_frame_adj(const Pose& P, float flag, const FrameAdj& gF, PoseAdj& gP) {
  gP.p += gF.o;
  gP.q.w += gF.q.w;
  gP.q.x += gF.q.x;
  gP.q.y += gF.q.y;
  gP.q.z += gF.q.z;
  float off = P.gap / 2.f * flag;
  float3 goff = f3(0, 0, 0);
  qrot_adj(P.q, f3(off, 0.f, 0.f), gF.o, gP.q, goff);
  gP.gap += goff.x * flag * 0.5f;
}
DSK_DEV void tool_frame_adj(const FrameAdj& gF, PoseAdj& gP) {
  gP.p += gF.o;
  gP.q.w += gF.q.w;
  gP.q.x += gF.q.x;
  gP.q.y += gF.q.y;
  gP.q.z += gF.q.z;
  gP.gap += gF.aux;   // non-zero for Chopsticks only
}

// Primitive.collide / Gripper.collide for one tool at one grid node
DSK_DEV float3 tool_collide(const ToolParams& T, const Pose& P0, const Pose& P1, float3 p, float3 v, float dt) {
  if (is_gripper(T.type)) {  // primitives.py:507-511 / :636-640: jaws applied sequentially
    const int jk = sdf_kind(T.type);
    v = contact_forward(T, jk, jaw_frame(P0, -1.f), jaw_frame(P1, -1.f), p, v, dt, true);
    v = contact_forward(T, jk, jaw_frame(P0, 1.f), jaw_frame(P1, 1.f), p, v, dt, true);
    return v;
  }
  return contact_forward(T, sdf_kind(T.type), tool_frame(P0), tool_frame(P1), p, v, dt, false);
}
// adjoint; v is the velocity ENTERING this tool
DSK_DEV float3 tool_collide_adj(const ToolParams& T, const Pose& P0, const Pose& P1, float3 p, float3 v, float dt,
                                float3 gout, PoseAdj& g0, PoseAdj& g1) {
  if (is_gripper(T.type)) {
    const int jk = sdf_kind(T.type);
    Frame a0 = jaw_frame(P0, -1.f), a1 = jaw_frame(P1, -1.f), b0 = jaw_frame(P0, 1.f), b1 = jaw_frame(P1, 1.f);
    float3 vmid = contact_forward(T, jk, a0, a1, p, v, dt, true);
    FrameAdj ga0 = frame_adj_zero(), ga1 = frame_adj_zero(), gb0 = frame_adj_zero(), gb1 = frame_adj_zero();
    float3 gmid = contact_adjoint(T, jk, b0, b1, p, vmid, dt, true, gout, gb0, gb1);
    float3 gin = contact_adjoint(T, jk, a0, a1, p, v, dt, true, gmid, ga0, ga1);
    jaw_frame_adj(P0, 1.f, gb0, g0);
    jaw_frame_adj(P1, 1.f, gb1, g1);
    jaw_frame_adj(P0, -1.f, ga0, g0);
    jaw_frame_adj(P1, -1.f, ga1, g1);
    return gin;
  }
  FrameAdj f0 = frame_adj_zero(), f1 = frame_adj_zero();
  float3 gin = contact_adjoint(T, sdf_kind(T.type), tool_frame(P0), tool_frame(P1), p, v, dt, false, gout, f0, f1);
  tool_frame_adj(f0, g0);
  tool_frame_adj(f1, g1);
  return gin;
}

// Primitive.sdf / Gripper.sdf and normals at tool level (collision projection, min-dist observation)
DSK_DEV float tool_sdf(const ToolParams& T, const Pose& P, float3 p) {
  if (is_gripper(T.type))
    return tmin(frame_sdf(T, sdf_kind(T.type), jaw_frame(P, -1.f), p), frame_sdf(T, sdf_kind(T.type), jaw_frame(P, 1.f), p));
  return frame_sdf(T, sdf_kind(T.type), tool_frame(P), p);
}
DSK_DEV void tool_sdf_adj(const ToolParams& T, const Pose& P, float3 p, float gd, PoseAdj& gP, float3& gp) {
  if (is_gripper(T.type)) {
    Frame a = jaw_frame(P, -1.f), b = jaw_frame(P, 1.f);
    float da = frame_sdf(T, sdf_kind(T.type), a, p), db = frame_sdf(T, sdf_kind(T.type), b, p);
    FrameAdj g = frame_adj_zero();
    if (da < db) {
      frame_sdf_adj(T, sdf_kind(T.type), a, p, gd, g, gp);
      jaw_frame_adj(P, -1.f, g, gP);
    } else {
      frame_sdf_adj(T, sdf_kind(T.type), b, p, gd, g, gp);
      jaw_frame_adj(P, 1.f, g, gP);
    }
    return;
  }
  FrameAdj g = frame_adj_zero();
  frame_sdf_adj(T, sdf_kind(T.type), tool_frame(P), p, gd, g, gp);
  tool_frame_adj(g, gP);
}
DSK_DEV float3 tool_normal(const ToolParams& T, const Pose& P, float3 p) {
  if (is_gripper(T.type)) {  // primitives.py:489-496
    Frame a = jaw_frame(P, -1.f), b = jaw_frame(P, 1.f);
    float da = frame_sdf(T, sdf_kind(T.type), a, p), db = frame_sdf(T, sdf_kind(T.type), b, p);
    return (da <= db) ? frame_normal(T, sdf_kind(T.type), a, p) : frame_normal(T, sdf_kind(T.type), b, p);
  }
  return frame_normal(T, sdf_kind(T.type), tool_frame(P), p);
}
DSK_DEV void tool_normal_adj(const ToolParams& T, const Pose& P, float3 p, float3 gN, PoseAdj& gP, float3& gp) {
  if (is_gripper(T.type)) {
    Frame a = jaw_frame(P, -1.f), b = jaw_frame(P, 1.f);
    float da = frame_sdf(T, sdf_kind(T.type), a, p), db = frame_sdf(T, sdf_kind(T.type), b, p);
    FrameAdj g = frame_adj_zero();
    if (da <= db) {
      frame_normal_adj(T, sdf_kind(T.type), a, p, gN, g, gp);
      jaw_frame_adj(P, -1.f, g, gP);
    } else {
      frame_normal_adj(T, sdf_kind(T.type), b, p, gN, g, gp);
      jaw_frame_adj(P, 1.f, g, gP);
    }
    return;
  }
  FrameAdj g = frame_adj_zero();
  frame_normal_adj(T, sdf_kind(T.type), tool_frame(P), p, gN, g, gp);
  tool_frame_adj(g, gP);
}

// ---- tool kinematics ---------------------------------------------------------------------------------------
// forward_kinematics of one tool, primive_base.py:152-156 / primitives.py:120-136 / :456-460
struct ToolVel {
  float3 v, w;
  float gap_vel;
};
// the axis-angle increments of a tool are the same for every substep of an env step (set_velocity,
// primive_base.py:260-268): their quaternions are built once per step
struct ToolRotInc {
  Q4 a, b;   // RollingPinExt: a = w2quat(0,-dth,0), b = w2quat(0,dw,0); others: a = w2quat(w)
};
// RollingPinExt slides along its rolling direction by w[0] on top of the roll dw * 0.03 (primitives.py:129); the plain
// RollingPin only rolls (primitives.py:110)
DSK_DEV float rollingpin_slide(const ToolParams& T, const ToolVel& u) { return T.type == DSK_TOOL_ROLLINGPIN_EXT ? u.w.x : 0.f; }
DSK_DEV ToolRotInc tool_rot_inc(const ToolParams& T, const ToolVel& u) {
  ToolRotInc r;
  if (is_rollingpin(T.type)) {
    r.a = w2quat(f3(0.f, -u.v.y, 0.f));
    r.b = w2quat(f3(0.f, u.v.x, 0.f));
  } else {
    r.a = w2quat(u.w);
    r.b = r.a;
  }
  return r;
}
DSK_DEV Pose tool_fk_inc(const ToolParams& T, const Pose& P, const ToolVel& u, const ToolRotInc& r) {
  Pose N;
  N.gap = P.gap;
  float3 step = u.v;
  if (is_rollingpin(T.type)) {
    float dw = u.v.x, dy = u.v.z;
    float3 y_dir = qrot_rn(P.q, f3(0.f, -1.f, 0.f));
    float3 cr = cross(f3(0.f, 1.f, 0.f), y_dir);
    float3 x_dir = T.type == DSK_TOOL_ROLLINGPIN_EXT ? (dw * 0.03f + u.w.x) * cr      // primitives.py:129
                                                     : 0.03f * (dw * cr);              // primitives.py:110
    x_dir.y = dy;
    N.q = qmul(r.a, qmul(P.q, r.b));
    step = x_dir;
  } else if (has_gap(T.type)) {   // Chopsticks (primitives.py:230-234) has no upper clamp: max_gap = +inf (fill_tool)
    N.gap = tmin(tmax(P.gap - u.gap_vel, T.min_gap), T.max_gap);
    N.q = qmul(P.q, r.a);
  } else {
    N.q = qmul(r.a, P.q);
  }
  N.p = f3(tmax(tmin(P.p.x + step.x, T.hi[0]), T.lo[0]), tmax(tmin(P.p.y + step.y, T.hi[1]), T.lo[1]),
           tmax(tmin(P.p.z + step.z, T.hi[2]), T.lo[2]));
  return N;
}
DSK_DEV Pose tool_fk(const ToolParams& T, const Pose& P, const ToolVel& u) {
  Pose N;
  N.gap = P.gap;
  float3 step = u.v;
  if (is_rollingpin(T.type)) {
    float dw = u.v.x, dth = u.v.y, dy = u.v.z;
    float3 y_dir = qrot_rn(P.q, f3(0.f, -1.f, 0.f));
    float3 cr = cross(f3(0.f, 1.f, 0.f), y_dir);
    float3 x_dir = T.type == DSK_TOOL_ROLLINGPIN_EXT ? (dw * 0.03f + u.w.x) * cr      // primitives.py:129
                                                     : 0.03f * (dw * cr);              // primitives.py:110
    x_dir.y = dy;
    N.q = qmul(w2quat(f3(0.f, -dth, 0.f)), qmul(P.q, w2quat(f3(0.f, dw, 0.f))));
    step = x_dir;
  } else if (has_gap(T.type)) {
    N.gap = tmin(tmax(P.gap - u.gap_vel, T.min_gap), T.max_gap);
    N.q = qmul(P.q, w2quat(u.w));
  } else {
    N.q = qmul(w2quat(u.w), P.q);
  }
  N.p = f3(tmax(tmin(P.p.x + step.x, T.hi[0]), T.lo[0]), tmax(tmin(P.p.y + step.y, T.hi[1]), T.lo[1]),
           tmax(tmin(P.p.z + step.z, T.hi[2]), T.lo[2]));
  return N;
}
// adjoint of tool_fk: given g(N) accumulates g(P), g(u)
DSK_DEV void tool_fk_adj(const ToolParams& T, const Pose& P, const ToolVel& u, const PoseAdj& gN, PoseAdj& gP,
                         ToolVel& gu) {
  float3 step = u.v;
  float3 y_dir = f3(0, 0, 0), cr = f3(0, 0, 0);
  float sc = 0.f;
  if (is_rollingpin(T.type)) {
    y_dir = qrot_rn(P.q, f3(0.f, -1.f, 0.f));
    cr = cross(f3(0.f, 1.f, 0.f), y_dir);
    sc = u.v.x * 0.03f + rollingpin_slide(T, u);
    step = sc * cr;
    step.y = u.v.z;
  }
  // position clamp: max(min(a, hi), lo)
  float3 gstep;
  {
    float a0 = P.p.x + step.x, a1 = P.p.y + step.y, a2 = P.p.z + step.z;
    float g0 = (a0 < T.hi[0] && T.lo[0] < tmin(a0, T.hi[0])) ? gN.p.x : 0.f;
    float g1 = (a1 < T.hi[1] && T.lo[1] < tmin(a1, T.hi[1])) ? gN.p.y : 0.f;
    float g2 = (a2 < T.hi[2] && T.lo[2] < tmin(a2, T.hi[2])) ? gN.p.z : 0.f;
    gstep = f3(g0, g1, g2);
    gP.p += gstep;
  }
  if (is_rollingpin(T.type)) {
    float dw = u.v.x, dth = u.v.y;
    // step = (sc*cr.x, dy, sc*cr.z)
    gu.v.z += gstep.y;
    float gsc = gstep.x * cr.x + gstep.z * cr.z;
    float3 gcr = f3(sc * gstep.x, 0.f, sc * gstep.z);
    gu.v.x += 0.03f * gsc;
    if (T.type == DSK_TOOL_ROLLINGPIN_EXT) gu.w.x += gsc;
    // cr = e_y x y_dir  -> g(y_dir) = gcr x e_y
    float3 gy = cross(gcr, f3(0.f, 1.f, 0.f));
    float3 gdummy = f3(0, 0, 0);
    qrot_adj(P.q, f3(0.f, -1.f, 0.f), gy, gP.q, gdummy);
    // N.q = qmul(qa, qmul(P.q, qb)), qa = w2quat(0,-dth,0), qb = w2quat(0,dw,0)
    Q4 qa = w2quat(f3(0.f, -dth, 0.f)), qb = w2quat(f3(0.f, dw, 0.f));
    Q4 inner = qmul(P.q, qb);
    Q4 gqa = {0, 0, 0, 0}, ginner = {0, 0, 0, 0}, gqb = {0, 0, 0, 0};
    qmul_adj(qa, inner, gN.q, gqa, ginner);
    qmul_adj(P.q, qb, ginner, gP.q, gqb);
    float3 ga = f3(0, 0, 0), gb = f3(0, 0, 0);
    w2quat_adj(f3(0.f, -dth, 0.f), gqa, ga);
    w2quat_adj(f3(0.f, dw, 0.f), gqb, gb);
    gu.v.y += -ga.y;
    gu.v.x += gb.y;
  } else if (has_gap(T.type)) {
    gu.v += gstep;
    float a = P.gap - u.gap_vel;
    float m1 = tmax(a, T.min_gap);
    float g = (m1 < T.max_gap) ? gN.gap : 0.f;     // min(m1, max_gap): m1 gets it iff m1 < max_gap
    g = (T.min_gap < a) ? g : 0.f;                 // max(a, min_gap): a gets it iff min_gap < a
    gP.gap += g;
    gu.gap_vel -= g;
    Q4 qw = w2quat(u.w);
    Q4 gqw = {0, 0, 0, 0};
    qmul_adj(P.q, qw, gN.q, gP.q, gqw);
    w2quat_adj(u.w, gqw, gu.w);
  } else {
    gu.v += gstep;
    Q4 qw = w2quat(u.w);
    Q4 gqw = {0, 0, 0, 0};
    qmul_adj(qw, P.q, gN.q, gqw, gP.q);
    w2quat_adj(u.w, gqw, gu.w);
  }
  if (!has_gap(T.type)) gP.gap += gN.gap;
}
DSK_DEV ToolVel action_to_vel(const ToolParams& T, const float* a, int S) {  // set_velocity, primive_base.py:260-268
  ToolVel u;
  u.v = f3(0, 0, 0);
  u.w = f3(0, 0, 0);
  u.gap_vel = 0.f;
  float fs = (float)S;
  if (T.action_dim > 0) {
    u.v = f3(a[0] * T.action_scale[0] / fs, a[1] * T.action_scale[1] / fs, a[2] * T.action_scale[2] / fs);
    if (T.action_dim > 3) u.w = f3(a[3] * T.action_scale[3] / fs, a[4] * T.action_scale[4] / fs, a[5] * T.action_scale[5] / fs);
    if (has_gap(T.type)) u.gap_vel = a[6] * T.action_scale[6] / fs;
  }
  return u;
}
