// TEST INFRASTRUCTURE: the handful of CUDA built-ins that mpm_math.cuh / tools.cuh use, for a plain g++ build
// (-ffp-contract=off, so the `_rn` intrinsics and ordinary float arithmetic both round once per operation).
#pragma once
#include <cmath>
#define DSK_DEV static inline
struct float3 {
  float x, y, z;
};
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
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
// fast-math intrinsics of the return map / SVD: exact host versions (the device ones are a few ulp off, which the tests'
// tolerances cover)
static inline float hc_logf(float a) { return std::log(a); }
static inline float hc_expf(float a) { return std::exp(a); }
#define __logf hc_logf   /* glibc declares its own __logf / __expf */
#define __expf hc_expf
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
