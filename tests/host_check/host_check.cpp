// TEST INFRASTRUCTURE: compiles the DEVICE tool math of the product (diffskill_b200/csrc/tools.cuh: SDFs, normals,
// contact response, forward kinematics and all their hand-derived adjoints) for the CPU and exposes it through a small
// C ABI, so that `-m "not gpu"` tests can compare it with the oracle's tape-AD restatement of the same reference code.
// The kernels themselves (launch geometry, shared memory, warp reductions) are only reachable on a GPU.
#define DSK_HOST_CHECK 1
#include <cstring>

#include "../../diffskill_b200/csrc/tools.cuh"

static ToolParams make_tool(const dsk_tool_desc& d) {  // mirrors fill_tool() of engine.cu (fp32 rounding of the doubles)
  ToolParams T;
  std::memset(&T, 0, sizeof T);
  T.type = d.type;
  T.action_dim = d.action_dim;
  for (int j = 0; j < 8; j++) T.action_scale[j] = (float)d.action_scale[j];
  T.friction = (float)d.friction;
  T.softness = (float)d.softness;
  for (int j = 0; j < 3; j++) {
    T.lo[j] = (float)d.lower_bound[j];
    T.hi[j] = (float)d.upper_bound[j];
    T.size[j] = (float)d.size[j];
  }
  T.h = (float)d.h;
  T.half_h = (float)(d.h / 2);
  T.r = (float)d.r;
  T.prism_h0 = (float)d.prism_h[0];
  T.prism_h1 = (float)d.prism_h[1];
  float w = (float)d.prot[0], x = (float)d.prot[1], y = (float)d.prot[2], z = (float)d.prot[3];
  T.prot = Q4{w, x, y, z};
  float n2 = w * w;
  n2 = n2 + x * x;
  n2 = n2 + y * y;
  n2 = n2 + z * z;
  float inv = 1.0f / std::sqrt(n2);
  T.prot_inv = Q4{inv * w, inv * -x, inv * -y, inv * -z};
  T.min_gap = (float)d.minimal_gap;
  T.max_gap = (float)d.maximal_gap;
  return T;
}
static void put_adj(float* o, const PoseAdj& g) {
  o[0] = g.p.x; o[1] = g.p.y; o[2] = g.p.z;
  o[3] = g.q.w; o[4] = g.q.x; o[5] = g.q.y; o[6] = g.q.z;
  o[7] = g.gap;
}

extern "C" {
float hc_tool_sdf(const dsk_tool_desc* d, const float* pose8, const float* p) {
  ToolParams T = make_tool(*d);
  return tool_sdf(T, load_pose(pose8), f3(p[0], p[1], p[2]));
}
void hc_tool_normal(const dsk_tool_desc* d, const float* pose8, const float* p, float* n) {
  ToolParams T = make_tool(*d);
  float3 r = tool_normal(T, load_pose(pose8), f3(p[0], p[1], p[2]));
  n[0] = r.x; n[1] = r.y; n[2] = r.z;
}
void hc_tool_collide(const dsk_tool_desc* d, const float* pose0, const float* pose1, const float* p, const float* v,
                     float dt, float* out) {
  ToolParams T = make_tool(*d);
  float3 r = tool_collide(T, load_pose(pose0), load_pose(pose1), f3(p[0], p[1], p[2]), f3(v[0], v[1], v[2]), dt);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// what: 0 sdf (gout[1]), 1 normal (gout[3]), 2 collide (gout[3]); out22 = [g(p) 3 | g(v_in) 3 | g(pose0) 8 | g(pose1) 8]
void hc_tool_probe_grad(const dsk_tool_desc* d, const float* pose0, const float* pose1, int what, const float* p,
                        const float* v, float dt, const float* gout, float* out22) {
  ToolParams T = make_tool(*d);
  Pose P0 = load_pose(pose0), P1 = load_pose(pose1);
  float3 pp = f3(p[0], p[1], p[2]), vv = f3(v[0], v[1], v[2]);
  PoseAdj g0 = pose_adj_zero(), g1 = pose_adj_zero();
  float3 gp = f3(0, 0, 0), gv = f3(0, 0, 0);
  if (what == 0) tool_sdf_adj(T, P0, pp, gout[0], g0, gp);
  else if (what == 1) tool_normal_adj(T, P0, pp, f3(gout[0], gout[1], gout[2]), g0, gp);
  else gv = tool_collide_adj(T, P0, P1, pp, vv, dt, f3(gout[0], gout[1], gout[2]), g0, g1);
  out22[0] = gp.x; out22[1] = gp.y; out22[2] = gp.z;
  out22[3] = gv.x; out22[4] = gv.y; out22[5] = gv.z;
  put_adj(out22 + 6, g0);
  put_adj(out22 + 14, g1);
}
// forward_kinematics from (state8, vel7 = v3 w3 gap_vel); gnext8 != NULL also the adjoint gout15 = [g(state) 8 | g(vel) 7]
void hc_tool_fk(const dsk_tool_desc* d, const float* state8, const float* vel7, float* next8, const float* gnext8,
                float* gout15) {
  ToolParams T = make_tool(*d);
  Pose P = load_pose(state8);
  ToolVel u;
  u.v = f3(vel7[0], vel7[1], vel7[2]);
  u.w = f3(vel7[3], vel7[4], vel7[5]);
  u.gap_vel = vel7[6];
  Pose N = tool_fk(T, P, u);
  Pose N2 = tool_fk_inc(T, P, u, tool_rot_inc(T, u));   // the per-step hoisted variant must agree
  next8[0] = N.p.x; next8[1] = N.p.y; next8[2] = N.p.z;
  next8[3] = N.q.w; next8[4] = N.q.x; next8[5] = N.q.y; next8[6] = N.q.z;
  next8[7] = N.gap;
  next8[8] = N2.p.x; next8[9] = N2.p.y; next8[10] = N2.p.z;
  next8[11] = N2.q.w; next8[12] = N2.q.x; next8[13] = N2.q.y; next8[14] = N2.q.z;
  next8[15] = N2.gap;
  if (!gnext8) return;
  PoseAdj gN = pose_adj_zero(), gP = pose_adj_zero();
  gN.p = f3(gnext8[0], gnext8[1], gnext8[2]);
  gN.q = Q4{gnext8[3], gnext8[4], gnext8[5], gnext8[6]};
  gN.gap = gnext8[7];
  ToolVel gu;
  gu.v = f3(0, 0, 0);
  gu.w = f3(0, 0, 0);
  gu.gap_vel = 0.f;
  tool_fk_adj(T, P, u, gN, gP, gu);
  put_adj(gout15, gP);
  gout15[8] = gu.v.x; gout15[9] = gu.v.y; gout15[10] = gu.v.z;
  gout15[11] = gu.w.x; gout15[12] = gu.w.y; gout15[13] = gu.w.z;
  gout15[14] = gu.gap_vel;
}
// action -> per-substep velocities (set_velocity, primive_base.py:260-268)
void hc_action_to_vel(const dsk_tool_desc* d, const float* action, int substeps, float* vel7) {
  ToolParams T = make_tool(*d);
  ToolVel u = action_to_vel(T, action, substeps);
  vel7[0] = u.v.x; vel7[1] = u.v.y; vel7[2] = u.v.z;
  vel7[3] = u.w.x; vel7[4] = u.w.y; vel7[5] = u.w.z;
  vel7[6] = u.gap_vel;
}
}
