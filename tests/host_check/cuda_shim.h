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
