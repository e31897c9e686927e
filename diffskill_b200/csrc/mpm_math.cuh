// Small register-resident 3-vector / quaternion / 3x3 helpers for the sm_100a kernels.
// `_rn` variants never contract to FMA: they are used wherever the reference's
// arithmetic is ill-conditioned (cell indexing, finite-difference SDF normals), so
// that the CUDA path and the CPU oracle round identically (DESIGN.md "Numerics").
#pragma once
#ifdef DSK_HOST_CHECK
// tests/host_check: the tool math (this header + tools.cuh) compiled by g++ for the CPU test suite, so that the device
// functions can be checked against the oracle without a GPU.  Never part of the product library.
#include "../../tests/host_check/cuda_shim.h"
#else
#include <cuda_runtime.h>

#define DSK_DEV __device__ __forceinline__
#endif

// Transcendentals of the return map and the reciprocals of the Jacobi SVD.  Product build: the hardware approximations
// (the reference runs ti.init(fast_math=True), plb/engine/taichi_env.py:20).  -DDSK_PRECISE_MATH (diagnostic library
// libdiffskill_mpm_pm.so, DSK_LIB=precise) swaps in the correctly rounded ones, to separate fast-math sensitivity of a
// scene's gradients (yield-surface branch flips) from defects.
#ifdef DSK_PRECISE_MATH
#define DSK_LOG(x) logf(x)
#define DSK_EXP(x) expf(x)
#define DSK_FDIV(a, b) ((a) / (b))
#define DSK_RSQRT(x) (1.f / sqrtf(x))
#else
#define DSK_LOG(x) __logf(x)
#define DSK_EXP(x) __expf(x)
#define DSK_FDIV(a, b) __fdividef(a, b)
#define DSK_RSQRT(x) rsqrtf(x)
#endif

struct Q4 {
  float w, x, y, z;
};
struct M3 {
  float m[9];  // row-major
};

// ---- exact-rounding scalar ops ------------------------------------------------------------
DSK_DEV float mul_rn(float a, float b) { return __fmul_rn(a, b); }
DSK_DEV float add_rn(float a, float b) { return __fadd_rn(a, b); }
DSK_DEV float sub_rn(float a, float b) { return __fsub_rn(a, b); }
// Taichi min/max tie rule is irrelevant for values; used for clarity only
DSK_DEV float tmin(float a, float b) { return a < b ? a : b; }
DSK_DEV float tmax(float a, float b) { return b < a ? a : b; }

// ---- float3 ---------------------------------------------------------------------------------
DSK_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
DSK_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
DSK_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
DSK_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
DSK_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
DSK_DEV float3 operator*(float3 a, float s) { return f3(s * a.x, s * a.y, s * a.z); }
DSK_DEV void operator+=(float3& a, float3 b) {
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
}
DSK_DEV void operator-=(float3& a, float3 b) {
  a.x -= b.x;
  a.y -= b.y;
  a.z -= b.z;
}
DSK_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DSK_DEV float3 cross(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
DSK_DEV float comp(float3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
DSK_DEV void setcomp(float3& a, int i, float v) {
  if (i == 0) a.x = v;
  else if (i == 1) a.y = v;
  else a.z = v;
}
// non-contracting versions (oracle op order: left-to-right sums)
DSK_DEV float3 sub3_rn(float3 a, float3 b) { return f3(sub_rn(a.x, b.x), sub_rn(a.y, b.y), sub_rn(a.z, b.z)); }
DSK_DEV float3 add3_rn(float3 a, float3 b) { return f3(add_rn(a.x, b.x), add_rn(a.y, b.y), add_rn(a.z, b.z)); }
DSK_DEV float dot_rn(float3 a, float3 b) {
  return add_rn(add_rn(mul_rn(a.x, b.x), mul_rn(a.y, b.y)), mul_rn(a.z, b.z));
}
DSK_DEV float3 cross_rn(float3 a, float3 b) {
  return f3(sub_rn(mul_rn(a.y, b.z), mul_rn(a.z, b.y)), sub_rn(mul_rn(a.z, b.x), mul_rn(a.x, b.z)),
            sub_rn(mul_rn(a.x, b.y), mul_rn(a.y, b.x)));
}

// ---- packed fp32x2 arithmetic: Blackwell FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add.rn.f32x2, sm_100+) -------------------------
// One instruction, two IEEE fp32 results (each rounded exactly like its scalar counterpart): the stencil loops and the
// Jacobi rotations work on (x, y) / (z, w) pairs, which halves their issue slots.  Host builds (tests/host_check) use the
// scalar operations.
#if defined(__CUDA_ARCH__)
DSK_DEV float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
DSK_DEV float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
DSK_DEV float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
#else
DSK_DEV float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
DSK_DEV float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
DSK_DEV float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#endif
DSK_DEV float2 f2(float x, float y) { return make_float2(x, y); }
DSK_DEV float2 bc2(float s) { return make_float2(s, s); }   // broadcast

// ---- quaternions (w,x,y,z), plb/engine/primitive/utils.py ------------------------------------
// qrot, utils.py:9-15
DSK_DEV float3 qrot_rn(Q4 q, float3 v) {
  float3 qv = f3(q.x, q.y, q.z);
  float3 uv = cross_rn(qv, v);
  float3 uuv = cross_rn(qv, uv);
  return f3(add_rn(v.x, mul_rn(2.f, add_rn(mul_rn(q.w, uv.x), uuv.x))),
            add_rn(v.y, mul_rn(2.f, add_rn(mul_rn(q.w, uv.y), uuv.y))),
            add_rn(v.z, mul_rn(2.f, add_rn(mul_rn(q.w, uv.z), uuv.z))));
}
// adjoint of out = qrot(q, v): accumulates into gq, gv
DSK_DEV void qrot_adj(Q4 q, float3 v, float3 go, Q4& gq, float3& gv) {
  float3 qv = f3(q.x, q.y, q.z);
  float3 uv = cross(qv, v);
  float3 o2 = 2.f * go;
  gv += go;
  gq.w += dot(o2, uv);
  float3 guv = q.w * o2;
  float3 guuv = o2;
  // uuv = qv x uv
  float3 gqv = cross(uv, guuv);
  guv += cross(guuv, qv);
  // uv = qv x v
  gqv += cross(v, guv);
  gv += cross(guv, qv);
  gq.x += gqv.x;
  gq.y += gqv.y;
  gq.z += gqv.z;
}
// Vector.normalized() of the conjugate: inv = 1/sqrt(dot); inv * c   (utils.py:53, primive_base.py:89-90)
DSK_DEV Q4 qconj_normalized_rn(Q4 r) {
  float n2 = add_rn(add_rn(add_rn(mul_rn(r.w, r.w), mul_rn(r.x, r.x)), mul_rn(r.y, r.y)), mul_rn(r.z, r.z));
  float inv = __fdiv_rn(1.f, __fsqrt_rn(n2));
  Q4 o;
  o.w = mul_rn(inv, r.w);
  o.x = mul_rn(inv, -r.x);
  o.y = mul_rn(inv, -r.y);
  o.z = mul_rn(inv, -r.z);
  return o;
}
// adjoint of n = conj(r)/|r| wrt r
DSK_DEV void qconj_normalized_adj(Q4 r, Q4 gn, Q4& gr) {
  float n2 = r.w * r.w + r.x * r.x + r.y * r.y + r.z * r.z;
  float inv = 1.f / sqrtf(n2);
  // c = conj(r); n = inv*c; inv = n2^-1/2
  float cw = r.w, cx = -r.x, cy = -r.y, cz = -r.z;
  float ginv = gn.w * cw + gn.x * cx + gn.y * cy + gn.z * cz;
  float gn2 = -0.5f * ginv * inv / n2;
  float gcw = inv * gn.w + 2.f * gn2 * cw;
  float gcx = inv * gn.x + 2.f * gn2 * cx;
  float gcy = inv * gn.y + 2.f * gn2 * cy;
  float gcz = inv * gn.z + 2.f * gn2 * cz;
  gr.w += gcw;
  gr.x -= gcx;
  gr.y -= gcy;
  gr.z -= gcz;
}
// qmul(q, r) = normalised Hamilton product q*r, utils.py:23-31
DSK_DEV Q4 qmul(Q4 q, Q4 r) {
  float w = r.w * q.w - r.x * q.x - r.y * q.y - r.z * q.z;
  float x = r.w * q.x + r.x * q.w - r.y * q.z + r.z * q.y;
  float y = r.w * q.y + r.x * q.z + r.y * q.w - r.z * q.x;
  float z = r.w * q.z - r.x * q.y + r.y * q.x + r.z * q.w;
  float n = sqrtf(w * w + x * x + y * y + z * z);
  Q4 o = {w / n, x / n, y / n, z / n};
  return o;
}
DSK_DEV void qmul_adj(Q4 q, Q4 r, Q4 go, Q4& gq, Q4& gr) {
  float w = r.w * q.w - r.x * q.x - r.y * q.y - r.z * q.z;
  float x = r.w * q.x + r.x * q.w - r.y * q.z + r.z * q.y;
  float y = r.w * q.y + r.x * q.z + r.y * q.w - r.z * q.x;
  float z = r.w * q.z - r.x * q.y + r.y * q.x + r.z * q.w;
  float n = sqrtf(w * w + x * x + y * y + z * z);
  // o = u/n : gu = go/n - u (u.go)/n^3
  float d = (w * go.w + x * go.x + y * go.y + z * go.z) / (n * n * n);
  float gw = go.w / n - w * d, gx = go.x / n - x * d, gy = go.y / n - y * d, gz = go.z / n - z * d;
  gq.w += gw * r.w + gx * r.x + gy * r.y + gz * r.z;
  gq.x += -gw * r.x + gx * r.w - gy * r.z + gz * r.y;
  gq.y += -gw * r.y + gx * r.z + gy * r.w - gz * r.x;
  gq.z += -gw * r.z - gx * r.y + gy * r.x + gz * r.w;
  gr.w += gw * q.w + gx * q.x + gy * q.y + gz * q.z;
  gr.x += -gw * q.x + gx * q.w + gy * q.z - gz * q.y;
  gr.y += -gw * q.y - gx * q.z + gy * q.w + gz * q.x;
  gr.z += -gw * q.z + gx * q.y - gy * q.x + gz * q.w;
}
// w2quat, utils.py:34-47
DSK_DEV Q4 w2quat(float3 aa) {
  float w = sqrtf(dot(aa, aa) + 1e-16f);
  Q4 o = {1.f, 0.f, 0.f, 0.f};
  if (w > 1e-9f) {
    float s, c;
    sincosf(w / 2.f, &s, &c);
    o.w = c;
    o.x = (aa.x / w) * s;
    o.y = (aa.y / w) * s;
    o.z = (aa.z / w) * s;
  }
  return o;
}
DSK_DEV void w2quat_adj(float3 aa, Q4 go, float3& gaa) {
  float w = sqrtf(dot(aa, aa) + 1e-16f);
  if (w > 1e-9f) {
    float s, c;
    sincosf(w / 2.f, &s, &c);
    float3 gv = f3(go.x, go.y, go.z);
    // o.w = cos(w/2); o.xyz = (aa/w)*s
    float gw = -0.5f * s * go.w;
    float gs = dot(gv, aa) / w;
    gw += 0.5f * c * gs;
    gw += -s * dot(gv, aa) / (w * w);
    gaa += (s / w) * gv;
    // w = sqrt(aa.aa + eps)
    gaa += (gw / w) * aa;
  }
}

// ---- 3x3 ------------------------------------------------------------------------------------
DSK_DEV M3 mm(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      C.m[i * 3 + j] = A.m[i * 3] * B.m[j] + A.m[i * 3 + 1] * B.m[3 + j] + A.m[i * 3 + 2] * B.m[6 + j];
  return C;
}
// A * B^T
DSK_DEV M3 mmT(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      C.m[i * 3 + j] = A.m[i * 3] * B.m[j * 3] + A.m[i * 3 + 1] * B.m[j * 3 + 1] + A.m[i * 3 + 2] * B.m[j * 3 + 2];
  return C;
}
// A^T * B
DSK_DEV M3 mTm(const M3& A, const M3& B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C.m[i * 3 + j] = A.m[i] * B.m[j] + A.m[3 + i] * B.m[3 + j] + A.m[6 + i] * B.m[6 + j];
  return C;
}
DSK_DEV float det3(const M3& a) {
  return a.m[0] * (a.m[4] * a.m[8] - a.m[7] * a.m[5]) - a.m[3] * (a.m[1] * a.m[8] - a.m[7] * a.m[2]) +
         a.m[6] * (a.m[1] * a.m[5] - a.m[4] * a.m[2]);
}
// cofactor matrix: d det / dA
DSK_DEV M3 cof3(const M3& a) {
  M3 c;
  c.m[0] = a.m[4] * a.m[8] - a.m[5] * a.m[7];
  c.m[1] = a.m[5] * a.m[6] - a.m[3] * a.m[8];
  c.m[2] = a.m[3] * a.m[7] - a.m[4] * a.m[6];
  c.m[3] = a.m[2] * a.m[7] - a.m[1] * a.m[8];
  c.m[4] = a.m[0] * a.m[8] - a.m[2] * a.m[6];
  c.m[5] = a.m[1] * a.m[6] - a.m[0] * a.m[7];
  c.m[6] = a.m[1] * a.m[5] - a.m[2] * a.m[4];
  c.m[7] = a.m[2] * a.m[3] - a.m[0] * a.m[5];
  c.m[8] = a.m[0] * a.m[4] - a.m[1] * a.m[3];
  return c;
}
DSK_DEV float3 mv(const M3& A, float3 x) {
  return f3(A.m[0] * x.x + A.m[1] * x.y + A.m[2] * x.z, A.m[3] * x.x + A.m[4] * x.y + A.m[5] * x.z,
            A.m[6] * x.x + A.m[7] * x.y + A.m[8] * x.z);
}
DSK_DEV float3 mTv(const M3& A, float3 x) {
  return f3(A.m[0] * x.x + A.m[3] * x.y + A.m[6] * x.z, A.m[1] * x.x + A.m[4] * x.y + A.m[7] * x.z,
            A.m[2] * x.x + A.m[5] * x.y + A.m[8] * x.z);
}

// warp helpers
#if !defined(DSK_HOST_CHECK) || defined(DSK_HOST_SIMT)
DSK_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
