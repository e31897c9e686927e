// TEST INFRASTRUCTURE: the handful of CUDA built-ins that mpm_math.cuh / tools.cuh use, for a plain g++ build
// (-ffp-contract=off, so the `_rn` intrinsics and ordinary float arithmetic both round once per operation).
#pragma once
#include <cmath>
#define DSK_DEV static inline
struct float3 {
  float x, y, z;
};
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
// sincosf: glibc <cmath> (GNU extension) has the same signature
struct float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
// fast-math intrinsics of the return map / SVD.  Default: exact host versions (the device ones are a few ulp off, which
// the tests' tolerances cover).  -DHC_APPROX: each carries a deterministic pseudo-random error of the size the hardware
// approximations are allowed (log: 2^-21.4 absolute; exp, div, rsqrt: 2 ulp) -- used to study how far a scene's gradients
// move under the fast-math freedom (scripts/fastmath_sensitivity.py), never by the tests.
#include <cstdint>
#include <cstring>
static inline float hc_noise(float x) {   // in [-1, 1], deterministic in the bits of x
  uint32_t h;
  std::memcpy(&h, &x, 4);
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return (float)(h & 0xFFFFFF) / (float)0x7FFFFF - 1.0f;
}
#ifdef HC_APPROX   /* bit mask: 1 log, 2 exp, 4 div + rsqrt (the Jacobi SVD) */
#define HC_APPROX_ON(bit) ((HC_APPROX) & (bit))
#else
#define HC_APPROX_ON(bit) 0
#endif
static inline float hc_logf(float a) {
  return HC_APPROX_ON(1) ? (float)(std::log((double)a) + 3.7e-7 * hc_noise(a)) : std::log(a);
}
static inline float hc_expf(float a) {
  return HC_APPROX_ON(2) ? (float)(std::exp((double)a) * (1.0 + 2.4e-7 * hc_noise(a))) : std::exp(a);
}
static inline float rsqrtf(float a) {
  return HC_APPROX_ON(4) ? (float)(1.0 / std::sqrt((double)a) * (1.0 + 2.4e-7 * hc_noise(a))) : 1.0f / std::sqrt(a);
}
static inline float __fdividef(float a, float b) {
  return HC_APPROX_ON(4) ? (float)((double)a / (double)b * (1.0 + 2.4e-7 * hc_noise(a + b))) : a / b;
}
#define __logf hc_logf   /* glibc declares its own __logf / __expf */
#define __expf hc_expf
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
#ifdef DSK_HOST_SIMT   /* warp / block primitives and the kernel-launch emulation */
#include "simt_shim.h"
#endif
#ifdef DSK_HOST_EMU    /* the CUDA runtime calls of engine.cu */
#include "cuda_rt_shim.h"
#endif
