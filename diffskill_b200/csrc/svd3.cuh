// Register-resident 3x3 SVD and its adjoint for the plasticity return map.
//
// Replaces the third-party `ti.svd` intrinsic (call site plb/engine/mpm_simulator.py:129)
// and `backward_svd` (mpm_simulator.py:131-156).  Convention (as ti.svd): U, V proper
// rotations, singular values ordered by decreasing magnitude, sign carried by the last one.
// Algorithm: 4 cyclic sweeps of one-sided (Hestenes) Jacobi on the columns of F, column
// sort, U from the normalised columns with u2 = u0 x u1.  Everything is fully unrolled
// so the 3x3s stay in registers.
#pragma once
#include "mpm_math.cuh"

DSK_DEV void jacobi_pair(float& b0p, float& b1p, float& b2p, float& b0q, float& b1q, float& b2q, float& v0p,
                         float& v1p, float& v2p, float& v0q, float& v1q, float& v2q) {
  float al = b0p * b0p + b1p * b1p + b2p * b2p;
  float be = b0q * b0q + b1q * b1q + b2q * b2q;
  float ga = b0p * b0q + b1p * b1q + b2p * b2q;
  if (ga != 0.f) {
    // approximate reciprocals (MUFU, <= 2 ulp): for the rotation ANGLE an error only changes how fast the sweeps converge;
    // zeta -> inf gives t -> 0
    float zeta = DSK_FDIV(be - al, 2.f * ga);
    float t = copysignf(1.f, zeta) * DSK_FDIV(1.f, fabsf(zeta) + sqrtf(1.f + zeta * zeta));
#ifndef DSK_BIASED_COSINE
    // The cosine is different: unless c^2 (1 + t^2) = 1 every rotation rescales the columns of V and B, and a BIAS of the
    // approximation (MUFU.RSQ; a plain fp32 Newton step has -1.4e-8 on average) makes sigma, R = U V^T and U f(S) V^T
    // drift the same way for every particle in every substep: v 1e-5 per substep on the B200, and the branch a soft-dough
    // scene's gradient sits on flips (Rope-v1: 3-step action gradient 3.5e-3 from the oracle).
    // The FMA-residual (Markstein) step below is unbiased to 3e-10 whatever the bias of its input, for 4 instructions.
    // On the CPU emulation of the engine under GPU-like arithmetic (2-ulp errors on every MUFU result, FMA contraction)
    // the biased cosine fails exactly the six Rope-v1 gradient cases the B200 failed, this one passes all 48
    // (profiles/r01j_cpu_emulated_gpu_like_arithmetic_*.log); -DDSK_BIASED_COSINE restores the old behaviour.
    float w = fmaf(t, t, 1.f);
    float y = DSK_RSQRT(w);
    float r = fmaf(-(w * y), 0.5f * y, 0.5f);   // 0.5 - (w y)(y / 2): the residual, one rounding
    float c = fmaf(y, r, y), s = c * t;
#else
    float c = DSK_RSQRT(1.f + t * t), s = c * t;
#endif
    float a, b;
    a = b0p; b = b0q; b0p = c * a - s * b; b0q = s * a + c * b;
    a = b1p; b = b1q; b1p = c * a - s * b; b1q = s * a + c * b;
    a = b2p; b = b2q; b2p = c * a - s * b; b2q = s * a + c * b;
    a = v0p; b = v0q; v0p = c * a - s * b; v0q = s * a + c * b;
    a = v1p; b = v1q; v1p = c * a - s * b; v1q = s * a + c * b;
    a = v2p; b = v2q; v2p = c * a - s * b; v2q = s * a + c * b;
  }
}

DSK_DEV void swapneg(bool doit, float& n2p, float& n2q, float& b0p, float& b1p, float& b2p, float& b0q, float& b1q,
                     float& b2q, float& v0p, float& v1p, float& v2p, float& v0q, float& v1q, float& v2q) {
  if (doit) {
    float t;
    t = b0p; b0p = b0q; b0q = -t;
    t = b1p; b1p = b1q; b1q = -t;
    t = b2p; b2p = b2q; b2q = -t;
    t = v0p; v0p = v0q; v0q = -t;
    t = v1p; v1p = v1q; v1q = -t;
    t = v2p; v2p = v2q; v2q = -t;
    t = n2p; n2p = n2q; n2q = t;
  }
}

// A = U diag(sig) V^T
DSK_DEV void svd3(const M3& A, M3& U, float3& sig, M3& V) {
  float b00 = A.m[0], b01 = A.m[1], b02 = A.m[2];
  float b10 = A.m[3], b11 = A.m[4], b12 = A.m[5];
  float b20 = A.m[6], b21 = A.m[7], b22 = A.m[8];
  float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
#pragma unroll
  for (int sw = 0; sw < 4; sw++) {
    jacobi_pair(b00, b10, b20, b01, b11, b21, v00, v10, v20, v01, v11, v21);  // (0,1)
    jacobi_pair(b00, b10, b20, b02, b12, b22, v00, v10, v20, v02, v12, v22);  // (0,2)
    jacobi_pair(b01, b11, b21, b02, b12, b22, v01, v11, v21, v02, v12, v22);  // (1,2)
  }
  float n0 = b00 * b00 + b10 * b10 + b20 * b20;
  float n1 = b01 * b01 + b11 * b11 + b21 * b21;
  float n2 = b02 * b02 + b12 * b12 + b22 * b22;
  swapneg(n0 < n1, n0, n1, b00, b10, b20, b01, b11, b21, v00, v10, v20, v01, v11, v21);
  swapneg(n0 < n2, n0, n2, b00, b10, b20, b02, b12, b22, v00, v10, v20, v02, v12, v22);
  swapneg(n1 < n2, n1, n2, b01, b11, b21, b02, b12, b22, v01, v11, v21, v02, v12, v22);
  float s0 = sqrtf(n0), s1 = sqrtf(n1);
  float3 u0, u1;
  if (s0 > 0.f) u0 = f3(b00 / s0, b10 / s0, b20 / s0);
  else u0 = f3(1.f, 0.f, 0.f);
  if (s1 > 1e-18f) {
    u1 = f3(b01 / s1, b11 / s1, b21 / s1);
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
  V.m[0] = v00; V.m[1] = v01; V.m[2] = v02;
  V.m[3] = v10; V.m[4] = v11; V.m[5] = v12;
  V.m[6] = v20; V.m[7] = v21; V.m[8] = v22;
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
