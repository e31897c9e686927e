// Register-resident 3x3 SVD and its adjoint for the plasticity return map.
//
// Replaces the third-party `ti.svd` intrinsic (call site plb/engine/mpm_simulator.py:129)
// and `backward_svd` (mpm_simulator.py:131-156).  Convention (as ti.svd): U, V proper
// rotations, singular values ordered by decreasing magnitude, sign carried by the last one.
// Algorithm: 4 cyclic sweeps of one-sided (Hestenes) Jacobi on the columns of F, column
// sort, U from the normalised columns with u2 = u0 x u1.  Everything is fully unrolled
// so the 3x3s stay in registers.  Round 2: the rotation costs two MUFU.RSQ instead of two
// reciprocals, a square root and a reciprocal square root, and the column updates are packed
// (FFMA2): 1 071 -> ~500 thread-instructions per particle (profiles/r02*_ncu_sections_*.md).
#pragma once
#include "mpm_math.cuh"

// unbiased reciprocal square root: MUFU.RSQ plus one FMA-residual (Markstein) step.  A BIAS of ~1e-8 in the cosine of the
// rotations rescales the columns of V and B the same way for every particle in every substep, sigma, R = U V^T and
// U f(S) V^T drift coherently (v 1e-5 per substep on the B200) and the branch a soft-dough scene's gradient sits on flips
// (Rope-v1: 3-step action gradient 3.5e-3 from the oracle).  The refined value is unbiased to 3e-10
// whatever the bias of the MUFU result.  -DDSK_BIASED_COSINE restores the raw MUFU value.
DSK_DEV float rsqrt_unbiased(float w) {
  float y = DSK_RSQRT(w);
#ifndef DSK_BIASED_COSINE
  float r = fmaf(-(w * y), 0.5f * y, 0.5f);   // 0.5 - (w y)(y / 2): the residual, one rounding
  y = fmaf(y, r, y);
#endif
  return y;
}

// One Hestenes rotation of columns p, q of B (and V).  A column lives in three packed pairs (b0, b1), (b2, v0), (v1, v2),
// so that the six 2x2 rotations are 12 packed instructions (FMUL2 / FFMA2).
// Rotation angle: tan(2 theta) = 2 ga / (be - al), |theta| <= pi/4, from two reciprocal square roots and no division:
//   h = sqrt(d^2 + (2 ga)^2),  cos^2(theta) = (1 + |d| / h) / 2,  sin(theta) = sign(d) (2 ga / h) / (2 cos(theta))
// (the same angle as t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = d / (2 ga), c = 1 / sqrt(1 + t^2), s = c t of the
// textbook form: tan(2 theta) = 2 t / (1 - t^2) = 1 / zeta).  c^2 + s^2 = 1 needs (d^2 + 4 ga^2) / h^2 = 1 and
// cos^2 / cos^2 = 1 without bias: both reciprocal square roots are the unbiased ones.
struct SvdCol {
  float2 a, b, c;   // (b0, b1), (b2, v0), (v1, v2)
};
DSK_DEV float col_dot(const SvdCol& p, const SvdCol& q) {
  float2 t = mul2(p.a, q.a);
  return fmaf(p.b.x, q.b.x, t.x + t.y);
}
DSK_DEV void jacobi_pair(SvdCol& p, SvdCol& q) {
  float al = col_dot(p, p), be = col_dot(q, q), ga = col_dot(p, q);
  float g2 = ga + ga, d = be - al;
  float n2 = fmaf(d, d, g2 * g2);
  if (n2 < 1e-30f) {   // al == be and a tiny ga: the squares underflow -- rescale (only the ratio d : g2 matters)
    d *= 1.8446744e19f;
    g2 *= 1.8446744e19f;
    n2 = fmaf(d, d, g2 * g2);
  }
  // converged pair: |ga| <= 6e-8 sqrt(al be), the rotation angle is below the rounding of the columns -- skipped.  Jacobi
  // converges quadratically, so after two sweeps most warps skip all of the remaining six rotations (a warp skips when all
  // its lanes do); rank-deficient columns (al be = 0) keep the exact test ga != 0.
  if (ga * ga > 4e-15f * (al * be) && n2 > 0.f) {
    float rh = rsqrt_unbiased(n2);
    float u = fmaf(0.5f * fabsf(d), rh, 0.5f);     // cos^2(theta), in [0.5, 1]
    float y = rsqrt_unbiased(u);
    float c = u * y;
    float s = (copysignf(0.5f, d) * y) * (g2 * rh);
    float2 c2 = bc2(c), s2 = bc2(s), n2 = bc2(-s);
    float2 t;
    t = fma2(c2, p.a, mul2(n2, q.a)); q.a = fma2(s2, p.a, mul2(c2, q.a)); p.a = t;
    t = fma2(c2, p.b, mul2(n2, q.b)); q.b = fma2(s2, p.b, mul2(c2, q.b)); p.b = t;
    t = fma2(c2, p.c, mul2(n2, q.c)); q.c = fma2(s2, p.c, mul2(c2, q.c)); p.c = t;
  }
}
// order two columns by decreasing norm; the swap (p, q) <- (q, -p) keeps V a proper rotation
DSK_DEV void swapneg(bool doit, float& n2p, float& n2q, SvdCol& p, SvdCol& q) {
  if (doit) {
    SvdCol t = p;
    p = q;
    q.a = f2(-t.a.x, -t.a.y);
    q.b = f2(-t.b.x, -t.b.y);
    q.c = f2(-t.c.x, -t.c.y);
    float n = n2p;
    n2p = n2q;
    n2q = n;
  }
}

// A = U diag(sig) V^T
DSK_DEV void svd3(const M3& A, M3& U, float3& sig, M3& V) {
  SvdCol c0 = {f2(A.m[0], A.m[3]), f2(A.m[6], 1.f), f2(0.f, 0.f)};
  SvdCol c1 = {f2(A.m[1], A.m[4]), f2(A.m[7], 0.f), f2(1.f, 0.f)};
  SvdCol c2 = {f2(A.m[2], A.m[5]), f2(A.m[8], 0.f), f2(0.f, 1.f)};
#pragma unroll
  for (int sw = 0; sw < 4; sw++) {
    jacobi_pair(c0, c1);  // (0,1)
    jacobi_pair(c0, c2);  // (0,2)
    jacobi_pair(c1, c2);  // (1,2)
  }
  float n0 = col_dot(c0, c0), n1 = col_dot(c1, c1), n2 = col_dot(c2, c2);
  swapneg(n0 < n1, n0, n1, c0, c1);
  swapneg(n0 < n2, n0, n2, c0, c2);
  swapneg(n1 < n2, n1, n2, c1, c2);
  float b00 = c0.a.x, b10 = c0.a.y, b20 = c0.b.x, b01 = c1.a.x, b11 = c1.a.y, b21 = c1.b.x;
  float b02 = c2.a.x, b12 = c2.a.y, b22 = c2.b.x;
  // sigma_i = |b_i| and u_i = b_i / |b_i| from one (unbiased) reciprocal square root each
  float r0 = n0 > 0.f ? rsqrt_unbiased(n0) : 0.f, r1 = n1 > 1e-36f ? rsqrt_unbiased(n1) : 0.f;
  float s0 = n0 * r0, s1 = fminf(n1 * r1, s0);   // n0 >= n1 after the sort; keep the order through the two roundings
  float3 u0, u1;
  if (n0 > 0.f) u0 = f3(b00 * r0, b10 * r0, b20 * r0);
  else u0 = f3(1.f, 0.f, 0.f);
  if (n1 > 1e-36f) {
    u1 = f3(b01 * r1, b11 * r1, b21 * r1);
  } else {  // rank <= 1: any unit vector orthogonal to u0
    float ax = fabsf(u0.x), ay = fabsf(u0.y), az = fabsf(u0.z);
    int k = ax < ay ? (ax < az ? 0 : 2) : (ay < az ? 1 : 2);
    float3 e = f3(k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f);
    float d = comp(u0, k);
    u1 = e - d * u0;
    u1 = (1.f / sqrtf(dot(u1, u1))) * u1;
  }
  float3 u2 = cross(u0, u1);
  sig = f3(s0, s1, u2.x * b02 + u2.y * b12 + u2.z * b22);
  U.m[0] = u0.x; U.m[1] = u1.x; U.m[2] = u2.x;
  U.m[3] = u0.y; U.m[4] = u1.y; U.m[5] = u2.y;
  U.m[6] = u0.z; U.m[7] = u1.z; U.m[8] = u2.z;
  V.m[0] = c0.b.y; V.m[1] = c1.b.y; V.m[2] = c2.b.y;
  V.m[3] = c0.c.x; V.m[4] = c1.c.x; V.m[5] = c2.c.x;
  V.m[6] = c0.c.y; V.m[7] = c1.c.y; V.m[8] = c2.c.y;
}

DSK_DEV float clamp_gap(float a) {  // mpm_simulator.py:184-192
  return a >= 0.f ? fmaxf(a, 1e-6f) : fminf(a, -1e-6f);
}

// backward_svd, mpm_simulator.py:136-156: returns the adjoint of F_tmp given gU, gsig (diagonal), gV.
// Written in the basis of (U,V): every term is U * M * V^T, so one sandwich at the end.
DSK_DEV M3 svd3_backward(const M3& gU, float3 gs, const M3& gV, const M3& U, float3 sg, const M3& V) {
  M3 a = mTm(U, gU);  // U^T gU
  M3 b = mTm(V, gV);  // V^T gV
  float s2[3] = {sg.x * sg.x, sg.y * sg.y, sg.z * sg.z};
  float s[3] = {sg.x, sg.y, sg.z};
  float gsd[3] = {gs.x, gs.y, gs.z};
  M3 Mx;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (i == j) {
        Mx.m[i * 3 + j] = gsd[i];
      } else {
        float Fij = 1.f / clamp_gap(s2[j] - s2[i]);
        // (F o (U^T gU - gU^T U)) Sigma  +  Sigma (F o (V^T gV - gV^T V))
        float A = Fij * (a.m[i * 3 + j] - a.m[j * 3 + i]);
        float B = Fij * (b.m[i * 3 + j] - b.m[j * 3 + i]);
        Mx.m[i * 3 + j] = A * s[j] + s[i] * B;
      }
    }
  return mmT(mm(U, Mx), V);
}
