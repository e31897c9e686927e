// Per-particle and per-node arithmetic of the substep kernels, free of launch geometry: constants, B-spline stencil,
// plasticity return map, the per-particle part of p2g and of p2g.grad (with svd_grad), the tail of g2p, the grid boundary
// conditions and the contact-response adjoint.  The kernels (kernels_fwd.cuh / kernels_bwd.cuh) call these; the CPU test
// suite compiles this header with g++ (tests/host_check, -DDSK_HOST_CHECK) and checks it against the oracle, so it must
// not use warp / block primitives.
#pragma once
#include "mpm_math.cuh"
#include "tools.cuh"
#include "svd3.cuh"

#define FRAME_COMPS 24
enum { CX = 0, CV = 3, CC = 6, CF = 15 };

struct SimConst {
  int n, nt, ntile, nnode;  // grid nodes per axis, tiles per axis, tiles, nodes (per env)
  int B, Npad, stride;      // envs, padded capacity, B*Npad
  int S, K, npairs;         // substeps, tools, tool-tool pairs
  int gf_mode;              // ground friction: 0 zero-normal, 1 Coulomb, 2 stick (mpm_simulator.py:245-258)
  float dt, dx, inv_dx, p_mass, c_stress, c_C, x_hi, x_lo, m_eps, ground_friction;
  float grav[3];            // (dt * g) * 30, mpm_simulator.py:235
  float mu, lam, ys;        // the scene's material; the kernels read these while no per-particle material was set (mat == null)
  int pairs[DSK_MAX_PAIRS][2];
#ifdef DSK_TIMELINE
  struct TlRec* tl;         // device timeline records (profiling build only)
  int tl_slot;              // record of this launch, < 0: none
#endif
};

// Quadratic B-spline stencil of one coordinate, mpm_simulator.py:201-204.  The cell index is
// integer work and must be bit-exact: separate roundings, truncation toward zero.
DSK_DEV void bspline1(float x, float inv_dx, int n, int& base, float& fx, float w[3]) {
  float xg = __fmul_rn(x, inv_dx);
  int b = (int)__fsub_rn(xg, 0.5f);
  b = max(0, min(b, n - 3));  // no-op for any state the reference can represent; keeps NaN/blown-up states in bounds
  base = b;
  fx = __fsub_rn(xg, (float)b);
  float a = 1.5f - fx, c = fx - 1.f, d = fx - 0.5f;
  w[0] = 0.5f * (a * a);
  w[1] = 0.75f - c * c;
  w[2] = 0.5f * (d * d);
}
// d w / d fx
DSK_DEV void bspline1_grad(float fx, float dw[3]) {
  dw[0] = -(1.5f - fx);
  dw[1] = -2.f * (fx - 1.f);
  dw[2] = fx - 0.5f;
}

struct Stencil {
  int ox[3], oy[3], oz[3];  // per-axis partial node offsets (tile-major)
  float wx[3], wy[3], wz[3];
  float fx, fy, fz;
  int bx, by, bz;
};
DSK_DEV void make_stencil(const SimConst& k, float x, float y, float z, Stencil& s) {
  bspline1(x, k.inv_dx, k.n, s.bx, s.fx, s.wx);
  bspline1(y, k.inv_dx, k.n, s.by, s.fy, s.wy);
  bspline1(z, k.inv_dx, k.n, s.bz, s.fz, s.wz);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    int X = s.bx + i, Y = s.by + i, Z = s.bz + i;
    s.ox[i] = ((X >> 2) * k.nt * k.nt << 6) | ((X & 3) << 4);
    s.oy[i] = ((Y >> 2) * k.nt << 6) | ((Y & 3) << 2);
    s.oz[i] = ((Z >> 2) << 6) | (Z & 3);
  }
}

// per-particle material (mu, lam, yield_stress): the sorted arrays, or the scene's constants while nobody called
// dsk_set_material (12 bytes per particle and kernel less)
DSK_DEV void load_mat(const SimConst& k, const float* __restrict__ mat, int g, float& mu, float& lam, float& ys) {
  if (mat) {
    mu = mat[g];
    lam = mat[k.stride + g];
    ys = mat[2 * k.stride + g];
  } else {
    mu = k.mu;
    lam = k.lam;
    ys = k.ys;
  }
}
DSK_DEV M3 load_m3(const float* __restrict__ f, int comp0, int stride, int gid) {
  M3 A;
#pragma unroll
  for (int i = 0; i < 9; i++) A.m[i] = f[(comp0 + i) * stride + gid];
  return A;
}
DSK_DEV void store_m3(float* __restrict__ f, int comp0, int stride, int gid, const M3& A) {
#pragma unroll
  for (int i = 0; i < 9; i++) f[(comp0 + i) * stride + gid] = A.m[i];
}
DSK_DEV float3 load_v3(const float* __restrict__ f, int comp0, int stride, int gid) {
  return f3(f[comp0 * stride + gid], f[(comp0 + 1) * stride + gid], f[(comp0 + 2) * stride + gid]);
}
DSK_DEV void store_v3(float* __restrict__ f, int comp0, int stride, int gid, float3 v) {
  f[comp0 * stride + gid] = v.x;
  f[(comp0 + 1) * stride + gid] = v.y;
  f[(comp0 + 2) * stride + gid] = v.z;
}

// ---- plasticity: compute_von_mises, mpm_simulator.py:165-182 ---------------------------------------
struct ReturnMap {
  bool yields;
  float3 sc;   // clamped sigma
  float3 eps;  // log sc
  float3 eh;   // deviatoric part
  float ehn;   // its eps-norm
  float dg;    // delta_gamma
  float3 e;    // exp of the returned log-strain
};
DSK_DEV M3 von_mises(const M3& Ftmp, const M3& U, float3 sig, const M3& V, float ys, float mu, ReturnMap& r) {
  r.sc = f3(tmax(sig.x, 0.05f), tmax(sig.y, 0.05f), tmax(sig.z, 0.05f));
  r.eps = f3(DSK_LOG(r.sc.x), DSK_LOG(r.sc.y), DSK_LOG(r.sc.z));   // MUFU log/exp: the reference runs fast_math=True
  float mean = (r.eps.x + r.eps.y + r.eps.z) / 3.f;
  r.eh = f3(r.eps.x - mean, r.eps.y - mean, r.eps.z - mean);
  r.ehn = sqrtf(dot(r.eh, r.eh) + 1e-8f);
  r.dg = r.ehn - ys / (2.f * mu);
  r.yields = r.dg > 0.f;
  if (r.yields) {
    float kf = r.dg / r.ehn;
    r.e = f3(DSK_EXP(r.eps.x - kf * r.eh.x), DSK_EXP(r.eps.y - kf * r.eh.y), DSK_EXP(r.eps.z - kf * r.eh.z));
    M3 US;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      US.m[i * 3 + 0] = U.m[i * 3 + 0] * r.e.x;
      US.m[i * 3 + 1] = U.m[i * 3 + 1] * r.e.y;
      US.m[i * 3 + 2] = U.m[i * 3 + 2] * r.e.z;
    }
    return mmT(US, V);
  }
  return Ftmp;
}

// everything p2g computes per particle before the scatter
struct P2GParticle {
  M3 Ftmp, U, V, newF, affine;
  float3 sig;
  ReturnMap rm;
  float J;
};
// HAVE_SVD: o.U, o.sig, o.V were read from the SVD tape of the forward pass (the adjoint's fused recompute,
// mpm_simulator.py:330-333, then skips the Jacobi sweeps -- a quarter of its instructions)
template <bool HAVE_SVD>
DSK_DEV void p2g_particle_impl(const SimConst& k, const M3& C, const M3& F, float mu, float lam, float ys, P2GParticle& o) {
  M3 Mx;
#pragma unroll
  for (int i = 0; i < 9; i++) Mx.m[i] = ((i % 4 == 0) ? 1.f : 0.f) + k.dt * C.m[i];
  o.Ftmp = mm(Mx, F);  // compute_F_tmp
  if (!HAVE_SVD) svd3(o.Ftmp, o.U, o.sig, o.V);
  o.newF = von_mises(o.Ftmp, o.U, o.sig, o.V, ys, mu, o.rm);
  o.J = det3(o.newF);
  M3 R = mmT(o.U, o.V);
  M3 A;
#pragma unroll
  for (int i = 0; i < 9; i++) A.m[i] = (2.f * mu) * (o.newF.m[i] - R.m[i]);
  M3 st = mmT(A, o.newF);
  float vol = (lam * o.J) * (o.J - 1.f);
  st.m[0] += vol;
  st.m[4] += vol;
  st.m[8] += vol;
#pragma unroll
  for (int i = 0; i < 9; i++) o.affine.m[i] = k.c_stress * st.m[i] + k.p_mass * C.m[i];
}
DSK_DEV void p2g_particle(const SimConst& k, const M3& C, const M3& F, float mu, float lam, float ys, P2GParticle& o) {
  p2g_particle_impl<false>(k, C, F, mu, lam, ys, o);
}
// SVD tape: [S][SVD_COMPS][stride] per step slot, U (9), sigma (3), V (9) of substep j's F_tmp
#define SVD_COMPS 21
DSK_DEV void store_svd(float* __restrict__ t, int stride, int gid, const P2GParticle& o) {
  store_m3(t, 0, stride, gid, o.U);
  store_v3(t, 9, stride, gid, o.sig);
  store_m3(t, 12, stride, gid, o.V);
}
// p2g_particle for the adjoint: from the tape when there is one
DSK_DEV void p2g_particle_adj(const SimConst& k, const float* __restrict__ svd, int gid, const M3& C, const M3& F, float mu,
                              float lam, float ys, P2GParticle& o) {
  if (svd) {
    o.U = load_m3(svd, 0, k.stride, gid);
    o.sig = load_v3(svd, 9, k.stride, gid);
    o.V = load_m3(svd, 12, k.stride, gid);
    p2g_particle_impl<true>(k, C, F, mu, lam, ys, o);
  } else {
    p2g_particle_impl<false>(k, C, F, mu, lam, ys, o);
  }
}

// ---- grid_op for one node ---------------------------------------------------------------------------
// boundary conditions of mpm_simulator.py:241-260; returns the post-boundary velocity
DSK_DEV float3 grid_boundary(const SimConst& k, int I0, int I1, int I2, float3 v) {
  const int bound = 3;
  int I[3] = {I0, I1, I2};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (I[d] < bound && comp(v, d) < 0.f) {
      if (d != 1 || k.gf_mode == 0) {
        setcomp(v, d, 0.f);
      } else if (k.gf_mode == 1) {
        float lin = v.y + 1e-30f;
        float3 off = f3((float)I0 * 1e-30f, (float)I1 * 1e-30f, (float)I2 * 1e-30f);
        float3 vit = f3(v.x - off.x, v.y - lin - off.y, v.z - off.z);
        float lit = sqrtf(dot(vit, vit) + 1e-8f);
        float sc = tmax(1.f + k.ground_friction * lin / lit, 0.f);
        v = f3(sc * (vit.x + off.x), 0.f, sc * (vit.z + off.z));
      } else {
        v = f3(0.f, 0.f, 0.f);
      }
    }
    if (I[d] > k.n - bound && comp(v, d) > 0.f) setcomp(v, d, 0.f);
  }
  return v;
}
DSK_DEV float3 grid_boundary_adj(const SimConst& k, int I0, int I1, int I2, float3 v, float3 g) {
  // replay forward, remember the inputs of the three stages, then reverse
  const int bound = 3;
  int I[3] = {I0, I1, I2};
  float3 vin[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    vin[d] = v;
    if (I[d] < bound && comp(v, d) < 0.f) {
      if (d != 1 || k.gf_mode == 0) {
        setcomp(v, d, 0.f);
      } else if (k.gf_mode == 1) {
        float lin = v.y + 1e-30f;
        float3 off = f3((float)I0 * 1e-30f, (float)I1 * 1e-30f, (float)I2 * 1e-30f);
        float3 vit = f3(v.x - off.x, v.y - lin - off.y, v.z - off.z);
        float lit = sqrtf(dot(vit, vit) + 1e-8f);
        float sc = tmax(1.f + k.ground_friction * lin / lit, 0.f);
        v = f3(sc * (vit.x + off.x), 0.f, sc * (vit.z + off.z));
      } else {
        v = f3(0.f, 0.f, 0.f);
      }
    }
    if (I[d] > k.n - bound && comp(v, d) > 0.f) setcomp(v, d, 0.f);
  }
#pragma unroll
  for (int d = 2; d >= 0; d--) {
    float3 u = vin[d];
    // state after the first `if` of stage d
    float3 mid = u;
    bool lower = I[d] < bound && comp(u, d) < 0.f;
    float lin = 0.f, lit = 1.f, a = 0.f, sc = 0.f;
    float3 vit = f3(0, 0, 0), off = f3(0, 0, 0);
    if (lower) {
      if (d != 1 || k.gf_mode == 0) {
        setcomp(mid, d, 0.f);
      } else if (k.gf_mode == 1) {
        lin = u.y + 1e-30f;
        off = f3((float)I0 * 1e-30f, (float)I1 * 1e-30f, (float)I2 * 1e-30f);
        vit = f3(u.x - off.x, u.y - lin - off.y, u.z - off.z);
        lit = sqrtf(dot(vit, vit) + 1e-8f);
        a = 1.f + k.ground_friction * lin / lit;
        sc = tmax(a, 0.f);
        mid = f3(sc * (vit.x + off.x), 0.f, sc * (vit.z + off.z));
      } else {
        mid = f3(0.f, 0.f, 0.f);
      }
    }
    if (I[d] > k.n - bound && comp(mid, d) > 0.f) setcomp(g, d, 0.f);
    if (lower) {
      if (d != 1 || k.gf_mode == 0) {
        setcomp(g, d, 0.f);
      } else if (k.gf_mode == 1) {
        g.y = 0.f;  // v_out[1] = 0
        float3 w = f3(vit.x + off.x, vit.y + off.y, vit.z + off.z);
        float gsc = dot(g, w);
        float3 gvit = sc * g;
        float ga = (0.f < a) ? gsc : 0.f;  // max(a, 0): a gets it iff 0 < a
        float glin = ga * k.ground_friction / lit;
        float glit = -ga * k.ground_friction * lin / (lit * lit);
        gvit += (glit / lit) * vit;
        // vit = u - lin*e_y - off ; lin = u.y + 1e-30
        glin -= gvit.y;
        g = gvit;
        g.y += glin;
      } else {
        g = f3(0.f, 0.f, 0.f);
      }
    }
  }
  return g;
}

// tail of g2p: new_C = c_C (M - new_v (x) fx) and the clamped position update
DSK_DEV void g2p_finish(const SimConst& k, const Stencil& s, float3 x, float3 nv, float3 m0, float3 m1, float3 m2,
                        float3& nx, M3& nC) {
  nC.m[0] = k.c_C * (m0.x - nv.x * s.fx); nC.m[1] = k.c_C * (m1.x - nv.x * s.fy); nC.m[2] = k.c_C * (m2.x - nv.x * s.fz);
  nC.m[3] = k.c_C * (m0.y - nv.y * s.fx); nC.m[4] = k.c_C * (m1.y - nv.y * s.fy); nC.m[5] = k.c_C * (m2.y - nv.y * s.fz);
  nC.m[6] = k.c_C * (m0.z - nv.z * s.fx); nC.m[7] = k.c_C * (m1.z - nv.z * s.fy); nC.m[8] = k.c_C * (m2.z - nv.z * s.fz);
  nx = f3(tmax(tmin(x.x + k.dt * nv.x, k.x_hi), k.x_lo), tmax(tmin(x.y + k.dt * nv.y, k.x_hi), k.x_lo),
          tmax(tmin(x.z + k.dt * nv.z, k.x_hi), k.x_lo));
}

// adjoint of contact_response given the geometry (D, cv, influence): returns g(v_in), outputs g(D), g(cv), g(influence)
DSK_DEV float3 contact_response_adj(float3 v, float3 D, float3 cv, float influence, float friction, bool eps14,
                                    float3 gout, float3& gD, float3& gcv, float& ginfl) {
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
  gcv = gout;
  float3 gu = (1.f - influence) * gout;
  ginfl = dot(gout, t2 - u);
  float3 gt2 = influence * gout;
  float3 gt = f3(0, 0, 0);
  float gnc = 0.f;
  if (flag) {
    float3 gq = mx * gt2;
    float gmx = dot(gt2, q);
    float ga2 = (a2 < 0.f) ? 0.f : gmx;
    float gtn = ga2 - dot(gq, t) / (tn * tn);
    gt += (1.f / tn) * gq;
    gnc += ga2 * friction;
    gt += (gtn / tn) * t;
  } else {
    gt += gt2;
  }
  gu += gt;
  float gmn = -dot(gt, D);
  gD = (-mn) * gt;
  if (nc < 0.f) gnc += gmn;
  gu += gnc * D;
  gD += gnc * u;
  gcv -= gu;
  return gu;
}

// everything of p2g.grad after the 27-node gather: S0 = sum w G, (m0,m1,m2) = columns of sum w G (x) offset,
// gw* = adjoints of the per-axis weights
DSK_DEV void p2g_adj_finish(const SimConst& k, int gid, const Stencil& s, const P2GParticle& o, float mu, float lam,
                            const M3& C, const M3& F, const float* __restrict__ adj_in, float* __restrict__ adj_out,
                            float3 S0, float3 m0, float3 m1, float3 m2, const float* gwx, const float* gwy,
                            const float* gwz) {
  float3 gv = k.p_mass * S0;
  float3 gf = (-k.dx) * mTv(o.affine, S0);
  M3 gA;  // adjoint of affine
  gA.m[0] = k.dx * (m0.x - S0.x * s.fx); gA.m[1] = k.dx * (m1.x - S0.x * s.fy); gA.m[2] = k.dx * (m2.x - S0.x * s.fz);
  gA.m[3] = k.dx * (m0.y - S0.y * s.fx); gA.m[4] = k.dx * (m1.y - S0.y * s.fy); gA.m[5] = k.dx * (m2.y - S0.y * s.fz);
  gA.m[6] = k.dx * (m0.z - S0.z * s.fx); gA.m[7] = k.dx * (m1.z - S0.z * s.fy); gA.m[8] = k.dx * (m2.z - S0.z * s.fz);
  float dw[3];
  bspline1_grad(s.fx, dw);
  gf.x += gwx[0] * dw[0] + gwx[1] * dw[1] + gwx[2] * dw[2];
  bspline1_grad(s.fy, dw);
  gf.y += gwy[0] * dw[0] + gwy[1] * dw[1] + gwy[2] * dw[2];
  bspline1_grad(s.fz, dw);
  gf.z += gwz[0] * dw[0] + gwz[1] * dw[1] + gwz[2] * dw[2];
  float3 gx = load_v3(adj_out, CX, k.stride, gid) + k.inv_dx * gf;

  // affine = c_stress * stress + p_mass * C
  M3 gCm, gS;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    gCm.m[i] = k.p_mass * gA.m[i];
    gS.m[i] = k.c_stress * gA.m[i];
  }
  // stress = A N^T + vol I,  A = 2mu (N - R),  vol = (lam J)(J - 1)
  M3 R = mmT(o.U, o.V);
  M3 A;
#pragma unroll
  for (int i = 0; i < 9; i++) A.m[i] = (2.f * mu) * (o.newF.m[i] - R.m[i]);
  M3 gAm = mm(gS, o.newF);   // g(A) = gS N
  M3 gN = mTm(gS, A);        // g(N) = gS^T A
  M3 gR;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    gN.m[i] += (2.f * mu) * gAm.m[i];
    gR.m[i] = -(2.f * mu) * gAm.m[i];
  }
  float gJ = lam * (2.f * o.J - 1.f) * (gS.m[0] + gS.m[4] + gS.m[8]);
  M3 cf = cof3(o.newF);
  M3 gFn = load_m3(adj_in, CF, k.stride, gid);  // F.grad[j+1]
#pragma unroll
  for (int i = 0; i < 9; i++) gN.m[i] += gJ * cf.m[i] + gFn.m[i];
  // R = U V^T
  M3 gU = mm(gR, o.V);
  M3 gV = mTm(gR, o.U);
  float3 gsig = f3(0, 0, 0);
  M3 gFt;
  if (!o.rm.yields) {
    gFt = gN;
  } else {
    const ReturnMap& r = o.rm;
    float e[3] = {r.e.x, r.e.y, r.e.z};
    // N = U diag(e) V^T
    M3 NV = mm(gN, o.V);     // gN V
    M3 NtU = mTm(gN, o.U);   // gN^T U
    M3 UtNV = mTm(o.U, NV);  // U^T gN V
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int q = 0; q < 3; q++) {
        gU.m[i * 3 + q] += NV.m[i * 3 + q] * e[q];
        gV.m[i * 3 + q] += NtU.m[i * 3 + q] * e[q];
      }
    float3 ge = f3(UtNV.m[0], UtNV.m[4], UtNV.m[8]);
    float3 gep = f3(ge.x * r.e.x, ge.y * r.e.y, ge.z * r.e.z);  // adjoint of the returned log strain
    float kf = r.dg / r.ehn;
    float3 geps = gep;
    float gk = -dot(gep, r.eh);
    float3 geh = (-kf) * gep;
    float gdg = gk / r.ehn;
    float gehn = -gk * r.dg / (r.ehn * r.ehn) + gdg;
    geh += (gehn / r.ehn) * r.eh;
    float gm = (geh.x + geh.y + geh.z) / 3.f;
    geps += f3(geh.x - gm, geh.y - gm, geh.z - gm);
    gsig = f3((0.05f < o.sig.x) ? geps.x / r.sc.x : 0.f, (0.05f < o.sig.y) ? geps.y / r.sc.y : 0.f,
              (0.05f < o.sig.z) ? geps.z / r.sc.z : 0.f);
#pragma unroll
    for (int i = 0; i < 9; i++) gFt.m[i] = 0.f;
  }
  M3 gsv = svd3_backward(gU, gsig, gV, o.U, o.sig, o.V);
#pragma unroll
  for (int i = 0; i < 9; i++) gFt.m[i] += gsv.m[i];
  // F_tmp = (I + dt C) F
  M3 Mx;
#pragma unroll
  for (int i = 0; i < 9; i++) Mx.m[i] = ((i % 4 == 0) ? 1.f : 0.f) + k.dt * C.m[i];
  M3 gM = mmT(gFt, F);
  M3 gF = mTm(Mx, gFt);
#pragma unroll
  for (int i = 0; i < 9; i++) gCm.m[i] += k.dt * gM.m[i];
  store_v3(adj_out, CX, k.stride, gid, gx);
  store_v3(adj_out, CV, k.stride, gid, gv);
  store_m3(adj_out, CC, k.stride, gid, gCm);
  store_m3(adj_out, CF, k.stride, gid, gF);
}

// ---- geometry of one contact (node, rigid frame): shared by grid_op and grid_op.grad -------------------------------
struct ContactGeom {   // per (node, frame), shared memory
  float influence;     // < 0: contact inactive
  float3 D, cv;
  float3 pl, nraw;     // tool-local node position, un-normalised local normal and its length: re-used by the adjoint
  float L;
};
DSK_DEV void contact_geometry(const ToolParams& T, int kind, const Frame& F0, const Frame& F1, float3 p, float dt,
                              ContactGeom& g) {
  float3 pl = inv_trans(F0, p);
  float dist;
  float3 pn = pl;     // point and shape the normal is taken at
  int nk = kind;
  if (kind == SDF_CHOPSTICKS) {   // primitives.py:245-261: min of the two sticks, the nearer one's normal
    ChopEval c = chop_eval(T, F0.aux, pl);
    dist = tmin(c.a, c.b);
    pn = c.a <= c.b ? c.pa : c.pb;
    nk = SDF_CAPSULE;
  } else {
    dist = local_sdf(T, kind, pl);
  }
  float influence;
  if (contact_active(dist, T.softness, influence)) {
    g.influence = influence;
    float L;
    float3 n = local_normal_raw(T, nk, pn, L);
    g.D = qrot_rn(F0.q, f3(__fdiv_rn(n.x, L), __fdiv_rn(n.y, L), __fdiv_rn(n.z, L)));   // primive_base.py:80-85
    // collider_v (primive_base.py:87-94): the relative position IS the local point
    float3 np = add3_rn(qrot_rn(F1.q, pl), F1.o);
    g.cv = f3(__fdiv_rn(sub_rn(np.x, p.x), dt), __fdiv_rn(sub_rn(np.y, p.y), dt), __fdiv_rn(sub_rn(np.z, p.z), dt));
    g.pl = pl;
    g.nraw = n;
    g.L = L;
  } else {
    g.influence = -1.f;
  }
}

// p2g.grad + svd_grad + compute_F_tmp.grad of one particle (the whole of k_p2g_adj): gathers the adjoints of
// (grid_v_in, grid_m) over the 27-node stencil, reads F.grad[j+1], writes x.grad (adding the g2p part already stored),
// v.grad, C.grad, F.grad of frame j
DSK_DEV void p2g_adj_particle(const SimConst& k, int gid, int env, const float* __restrict__ fin,
                              const float* __restrict__ adj_in, float* __restrict__ adj_out, const float* __restrict__ mat,
                              const float4* __restrict__ Ga, const float* __restrict__ svd_in) {
  float3 x = load_v3(fin, CX, k.stride, gid);
  float3 v = load_v3(fin, CV, k.stride, gid);
  M3 C = load_m3(fin, CC, k.stride, gid);
  M3 F = load_m3(fin, CF, k.stride, gid);
  float mu, lam, ys;
  load_mat(k, mat, gid, mu, lam, ys);
  P2GParticle o;
  p2g_particle_adj(k, svd_in, gid, C, F, mu, lam, ys, o);
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  const float4* Gae = Ga + (size_t)env * k.nnode;
  // contribution(node) = w * (a0 + i ax + j ay + l az, p_mass)  (see k_p2g); with S0 = sum w G and M = sum w G (x) offset:
  //   g(v) = p_mass S0 ; g(affine) = dx (M - S0 (x) fx) ; g(fx) through dpos = -dx affine^T S0 ; g(w) = G . a + gm p_mass
  float3 fxv = f3(s.fx, s.fy, s.fz);
  float3 a0 = k.p_mass * v - k.dx * mv(o.affine, fxv);
  float3 ax = f3(k.dx * o.affine.m[0], k.dx * o.affine.m[3], k.dx * o.affine.m[6]);
  float3 ay = f3(k.dx * o.affine.m[1], k.dx * o.affine.m[4], k.dx * o.affine.m[7]);
  float3 az = f3(k.dx * o.affine.m[2], k.dx * o.affine.m[5], k.dx * o.affine.m[8]);
  float gwx[3] = {0, 0, 0}, gwy[3] = {0, 0, 0}, gwz[3] = {0, 0, 0};
  float3 S0, m0, m1, m2;
  {
    // moments of the grid adjoint (S0 = sum w G, M = sum w G (x) offset) by sum factorisation along l, j, i as in
    // g2p_particle; the weight adjoints gw = G . (a, p_mass) per node with the affine value a(i, j, l) built
    // incrementally, (x, y) and (z, mass) as packed pairs
    const float2 wz0 = bc2(s.wz[0]), wz1 = bc2(s.wz[1]), wz2 = bc2(s.wz[2]), wz22 = bc2(2.f * s.wz[2]);
    const float2 axl = f2(ax.x, ax.y), ayl = f2(ay.x, ay.y), azl = f2(az.x, az.y);
    float2 S0p = f2(0.f, 0.f), m0p = f2(0.f, 0.f), m1p = f2(0.f, 0.f), m2p = f2(0.f, 0.f);
    float S0z = 0.f, m0z = 0.f, m1z = 0.f, m2z = 0.f;
    float2 ai_lo = f2(a0.x, a0.y);
    float ai_z = a0.z;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float2 Q = f2(0.f, 0.f), Qy = f2(0.f, 0.f), Ql = f2(0.f, 0.f);
      float Qz = 0.f, Qyz = 0.f, Qlz = 0.f;
      float2 aij_lo = ai_lo;
      float aij_z = ai_z;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int ob = s.ox[i] + s.oy[j];
        float4 g0 = Gae[ob + s.oz[0]], g1 = Gae[ob + s.oz[1]], g2 = Gae[ob + s.oz[2]];
        float2 R = fma2(wz2, f2(g2.x, g2.y), fma2(wz1, f2(g1.x, g1.y), mul2(wz0, f2(g0.x, g0.y))));
        float Rz = fmaf(s.wz[2], g2.z, fmaf(s.wz[1], g1.z, s.wz[0] * g0.z));
        float2 R1 = fma2(wz22, f2(g2.x, g2.y), mul2(wz1, f2(g1.x, g1.y)));
        float R1z = fmaf(wz22.x, g2.z, s.wz[1] * g1.z);
        const float wy = s.wy[j];
        Q = fma2(bc2(wy), R, Q);
        Qz = fmaf(wy, Rz, Qz);
        Ql = fma2(bc2(wy), R1, Ql);
        Qlz = fmaf(wy, R1z, Qlz);
        if (j) {
          Qy = fma2(bc2((float)j * wy), R, Qy);
          Qyz = fmaf((float)j * wy, Rz, Qyz);
        }
        // gw(i,j,l) = G . a(i,j,l) + gm p_mass
        const float wxy = s.wx[i] * wy;
        float2 t0 = fma2(f2(g0.z, g0.w), f2(aij_z, k.p_mass), mul2(f2(g0.x, g0.y), aij_lo));
        float2 a1 = add2(aij_lo, azl);
        float2 t1 = fma2(f2(g1.z, g1.w), f2(aij_z + az.z, k.p_mass), mul2(f2(g1.x, g1.y), a1));
        float2 a2 = add2(a1, azl);
        float2 t2 = fma2(f2(g2.z, g2.w), f2(aij_z + 2.f * az.z, k.p_mass), mul2(f2(g2.x, g2.y), a2));
        float gw0 = t0.x + t0.y, gw1 = t1.x + t1.y, gw2 = t2.x + t2.y;
        gwz[0] = fmaf(gw0, wxy, gwz[0]);
        gwz[1] = fmaf(gw1, wxy, gwz[1]);
        gwz[2] = fmaf(gw2, wxy, gwz[2]);
        float h = fmaf(gw2, s.wz[2], fmaf(gw1, s.wz[1], gw0 * s.wz[0]));
        gwx[i] = fmaf(h, wy, gwx[i]);
        gwy[j] = fmaf(h, s.wx[i], gwy[j]);
        aij_lo = add2(aij_lo, ayl);
        aij_z += ay.z;
      }
      const float wx = s.wx[i];
      S0p = fma2(bc2(wx), Q, S0p);   S0z = fmaf(wx, Qz, S0z);
      m1p = fma2(bc2(wx), Qy, m1p);  m1z = fmaf(wx, Qyz, m1z);
      m2p = fma2(bc2(wx), Ql, m2p);  m2z = fmaf(wx, Qlz, m2z);
      if (i) {
        m0p = fma2(bc2((float)i * wx), Q, m0p);
        m0z = fmaf((float)i * wx, Qz, m0z);
      }
      ai_lo = add2(ai_lo, axl);
      ai_z += ax.z;
    }
    S0 = f3(S0p.x, S0p.y, S0z);
    m0 = f3(m0p.x, m0p.y, m0z);
    m1 = f3(m1p.x, m1p.y, m1z);
    m2 = f3(m2p.x, m2p.y, m2z);
  }
  p2g_adj_finish(k, gid, s, o, mu, lam, C, F, adj_in, adj_out, S0, m0, m1, m2, gwx, gwy, gwz);
}

// ---- g2p.grad of one particle, around the scatter of the grid_v_out adjoints ------------------------------------------
struct G2PAdj {
  float3 gx;              // adjoint of x through the clamped position update
  float3 b0, cx, cy, cz;  // adjoint of grid_v_out[node(i,j,l)] = w * (b0 + i cx + j cy + l cz)
};
DSK_DEV void g2p_adj_begin(const SimConst& k, float3 x, float3 xn, float3 gxn, float3 gvn, const M3& gC, Stencil& s,
                           G2PAdj& c) {
  // x' = max(min(x + dt v', hi), lo): the adjoint passes iff lo < x' < hi
  float3 gt = f3((k.x_lo < xn.x && xn.x < k.x_hi) ? gxn.x : 0.f, (k.x_lo < xn.y && xn.y < k.x_hi) ? gxn.y : 0.f,
                 (k.x_lo < xn.z && xn.z < k.x_hi) ? gxn.z : 0.f);
  gvn += k.dt * gt;
  c.gx = gt;
  make_stencil(k, x.x, x.y, x.z, s);
  // adjoint of grid_v_out[node] = w * (gvn + c_C gC (offset - fx)) = w * (b0 + i cx + j cy + l cz)
  c.b0 = gvn - k.c_C * mv(gC, f3(s.fx, s.fy, s.fz));
  c.cx = f3(k.c_C * gC.m[0], k.c_C * gC.m[3], k.c_C * gC.m[6]);
  c.cy = f3(k.c_C * gC.m[1], k.c_C * gC.m[4], k.c_C * gC.m[7]);
  c.cz = f3(k.c_C * gC.m[2], k.c_C * gC.m[5], k.c_C * gC.m[8]);
}
DSK_DEV float4 g2p_adj_node(const Stencil& s, const G2PAdj& c, int i, int j, int l) {
  float w = s.wx[i] * s.wy[j] * s.wz[l];
  float3 a = c.b0 + (float)i * c.cx + (float)j * c.cy + (float)l * c.cz;
  return make_float4(w * a.x, w * a.y, w * a.z, 0.f);
}
// adjoint of the weights: d/dw [ g . (gvn + c_C gC dpos) ] = g . (b0 + i cx + j cy + l cz); adjoint of fx through dpos:
// -c_C gC^T (sum w g).  Gve: grid_v_out of the particle's env (tile-major).  Returns x.grad.
DSK_DEV float3 g2p_adj_finish(const SimConst& k, const Stencil& s, const float4* __restrict__ Gve, const M3& gC,
                              const G2PAdj& c) {
  float gwx[3] = {0, 0, 0}, gwy[3] = {0, 0, 0}, gwz[3] = {0, 0, 0};
  float3 sg;
  {
    // sg = sum w g per node; gw = g . a(i,j,l) with the affine value built incrementally ((x, y) packed, z scalar)
    const float2 cxl = f2(c.cx.x, c.cx.y), cyl = f2(c.cy.x, c.cy.y), czl = f2(c.cz.x, c.cz.y);
    float2 sgp = f2(0.f, 0.f);
    float sgz = 0.f;
    float2 ai_lo = f2(c.b0.x, c.b0.y);
    float ai_z = c.b0.z;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float2 aij_lo = ai_lo;
      float aij_z = ai_z;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int ob = s.ox[i] + s.oy[j];
        float4 g0 = Gve[ob + s.oz[0]], g1 = Gve[ob + s.oz[1]], g2 = Gve[ob + s.oz[2]];
        const float wy = s.wy[j], wxy = s.wx[i] * wy;
        float2 R = fma2(bc2(s.wz[2]), f2(g2.x, g2.y), fma2(bc2(s.wz[1]), f2(g1.x, g1.y), mul2(bc2(s.wz[0]), f2(g0.x, g0.y))));
        float Rz = fmaf(s.wz[2], g2.z, fmaf(s.wz[1], g1.z, s.wz[0] * g0.z));
        sgp = fma2(bc2(wxy), R, sgp);
        sgz = fmaf(wxy, Rz, sgz);
        float2 t0 = mul2(f2(g0.x, g0.y), aij_lo);
        float2 a1 = add2(aij_lo, czl);
        float2 t1 = mul2(f2(g1.x, g1.y), a1);
        float2 a2 = add2(a1, czl);
        float2 t2 = mul2(f2(g2.x, g2.y), a2);
        float gw0 = fmaf(g0.z, aij_z, t0.x + t0.y);
        float gw1 = fmaf(g1.z, aij_z + c.cz.z, t1.x + t1.y);
        float gw2 = fmaf(g2.z, aij_z + 2.f * c.cz.z, t2.x + t2.y);
        gwz[0] = fmaf(gw0, wxy, gwz[0]);
        gwz[1] = fmaf(gw1, wxy, gwz[1]);
        gwz[2] = fmaf(gw2, wxy, gwz[2]);
        float h = fmaf(gw2, s.wz[2], fmaf(gw1, s.wz[1], gw0 * s.wz[0]));
        gwx[i] = fmaf(h, wy, gwx[i]);
        gwy[j] = fmaf(h, s.wx[i], gwy[j]);
        aij_lo = add2(aij_lo, cyl);
        aij_z += c.cy.z;
      }
      ai_lo = add2(ai_lo, cxl);
      ai_z += c.cx.z;
    }
    sg = f3(sgp.x, sgp.y, sgz);
  }
  float3 gf = (-k.c_C) * mTv(gC, sg);
  float dw[3];
  bspline1_grad(s.fx, dw);
  gf.x += gwx[0] * dw[0] + gwx[1] * dw[1] + gwx[2] * dw[2];
  bspline1_grad(s.fy, dw);
  gf.y += gwy[0] * dw[0] + gwy[1] * dw[1] + gwy[2] * dw[2];
  bspline1_grad(s.fz, dw);
  gf.z += gwz[0] * dw[0] + gwz[1] * dw[1] + gwz[2] * dw[2];
  return c.gx + k.inv_dx * gf;
}

// ---- contact frames of the tools and the per-tile culling of the grid kernels ---------------------------------------
#define MAX_FRAMES 8
struct FrameTable {   // built once per CTA
  int n;              // number of contact frames
  int tool[MAX_FRAMES];
  float flag[MAX_FRAMES];   // -1 / +1 gripper jaw, 0 plain tool
};
DSK_DEV void build_frame_table(const SimConst& k, const ToolParams* sT, FrameTable& ft) {
  int n = 0;
  for (int t = 0; t < k.K; t++) {
    if (is_gripper(sT[t].type)) {
      ft.tool[n] = t; ft.flag[n++] = -1.f;
      ft.tool[n] = t; ft.flag[n++] = 1.f;
    } else {
      ft.tool[n] = t; ft.flag[n++] = 0.f;
    }
  }
  ft.n = n;
}
// tool parameters + frame table, passed to the grid kernels by value (constant bank)
struct GridTools {
  ToolParams T[DSK_MAX_TOOLS];
  FrameTable ft;
};
DSK_DEV Frame frame_of_pose(const Pose& P, float flag) { return flag == 0.f ? tool_frame(P) : jaw_frame(P, flag); }
struct TileFrames {    // per tile-iteration, shared memory
  Frame F0[MAX_FRAMES], F1[MAX_FRAMES];
  int active[MAX_FRAMES];
};

DSK_DEV bool tile_any_active(const TileFrames& tf, int n) {   // uniform over the CTA
  int a = 0;
  for (int f = 0; f < n; f++) a |= tf.active[f];
  return a != 0;
}
// prepares contact frame y of the tile's env (poses at substeps j and j+1) and decides whether the tile can touch
// it at all
DSK_DEV int prepare_frame(const SimConst& k, const ToolParams* sT, const FrameTable& ft, int y,
                          const float* __restrict__ poses, int env, int j, int tx, int ty, int tz, Frame& F0, Frame& F1) {
  int t = ft.tool[y];
  const float* a = poses + ((size_t)(env * (k.S + 1) + j) * k.K + t) * 8;
  Pose P0 = load_pose(a), P1 = load_pose(a + (size_t)k.K * 8);
  F0 = frame_of_pose(P0, ft.flag[y]);
  F1 = frame_of_pose(P1, ft.flag[y]);
  const ToolParams& T = sT[t];
  int kind = sdf_kind(T.type);
  float3 c = f3(((float)(tx * 4) + 1.5f) * k.dx, ((float)(ty * 4) + 1.5f) * k.dx, ((float)(tz * 4) + 1.5f) * k.dx);
  float reach = 2.6f * k.dx * 1.02f + 1e-4f + (T.softness > 0.f ? 2.302586f / T.softness : 0.f);
  // cheap bounding-sphere test first, exact SDF only for tiles near the tool.  All SDFs here are 1-Lipschitz:
  // beyond `reach` no node of the tile has dist <= 0 or influence > 0.1
  float3 dc = c - F0.o;
  float far = T.bound_r + reach;
  int act = dot(dc, dc) <= far * far;
  if (act) act = frame_sdf(T, kind, F0, c) <= reach;
  return act;
}

// g2p of one particle (mpm_simulator.py:264-283): new_v = sum w g ; new_C = 4 inv_dx sum w g (x) (offset - fx)
//   = c_C (M - new_v (x) fx),  M = sum w g (x) offset
DSK_DEV void g2p_particle(const SimConst& k, const Stencil& s, const float4* __restrict__ Ge, float3 x, float3& nx,
                          float3& nv, M3& nC) {
  // Sum factorisation over the tensor-product weights (round 2): first along l (R = sum_l wz_l g, R1 = sum_l l wz_l g per
  // (i, j)), then along j, then along i -- 160 packed / scalar multiply-adds instead of 27 x 12.  (x, y) travel as one
  // packed pair (FFMA2), z as a scalar.
  const float2 wz0 = bc2(s.wz[0]), wz1 = bc2(s.wz[1]), wz2 = bc2(s.wz[2]), wz22 = bc2(2.f * s.wz[2]);
  float2 nvp = f2(0.f, 0.f), m0p = f2(0.f, 0.f), m1p = f2(0.f, 0.f), m2p = f2(0.f, 0.f);
  float nvz = 0.f, m0z = 0.f, m1z = 0.f, m2z = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float2 Q = f2(0.f, 0.f), Qy = f2(0.f, 0.f), Ql = f2(0.f, 0.f);
    float Qz = 0.f, Qyz = 0.f, Qlz = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int o = s.ox[i] + s.oy[j];
      float4 g0 = Ge[o + s.oz[0]], g1 = Ge[o + s.oz[1]], g2 = Ge[o + s.oz[2]];
      float2 R = fma2(wz2, f2(g2.x, g2.y), fma2(wz1, f2(g1.x, g1.y), mul2(wz0, f2(g0.x, g0.y))));
      float Rz = fmaf(s.wz[2], g2.z, fmaf(s.wz[1], g1.z, s.wz[0] * g0.z));
      float2 R1 = fma2(wz22, f2(g2.x, g2.y), mul2(wz1, f2(g1.x, g1.y)));
      float R1z = fmaf(wz22.x, g2.z, s.wz[1] * g1.z);
      const float wy = s.wy[j];
      Q = fma2(bc2(wy), R, Q);
      Qz = fmaf(wy, Rz, Qz);
      Ql = fma2(bc2(wy), R1, Ql);
      Qlz = fmaf(wy, R1z, Qlz);
      if (j) {
        Qy = fma2(bc2((float)j * wy), R, Qy);
        Qyz = fmaf((float)j * wy, Rz, Qyz);
      }
    }
    const float wx = s.wx[i];
    nvp = fma2(bc2(wx), Q, nvp);   nvz = fmaf(wx, Qz, nvz);
    m1p = fma2(bc2(wx), Qy, m1p);  m1z = fmaf(wx, Qyz, m1z);
    m2p = fma2(bc2(wx), Ql, m2p);  m2z = fmaf(wx, Qlz, m2z);
    if (i) {
      m0p = fma2(bc2((float)i * wx), Q, m0p);
      m0z = fmaf((float)i * wx, Qz, m0z);
    }
  }
  nv = f3(nvp.x, nvp.y, nvz);
  g2p_finish(k, s, x, nv, f3(m0p.x, m0p.y, m0z), f3(m1p.x, m1p.y, m1z), f3(m2p.x, m2p.y, m2z), nx, nC);
}
