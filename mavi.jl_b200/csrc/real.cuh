// real.cuh — the arithmetic type of one build of the device code.  Every source of libmavi_cuda.so is compiled twice:
// once with real = double (namespace mavi_f64, the default Float64 mode) and once with -DMAVI_REAL_F32
// (real = float, namespace mavi_f32, the optional Float32 mode: states whose element type is Float32,
// src/init_states.jl:34,63 NUM_T).  capi.cu holds the extern "C" entry points and dispatches on MaviParams.dtype.
#pragma once

#include <cuda_runtime.h>

#ifdef MAVI_REAL_F32
#define MAVI_NS mavi_f32
#define MAVI_REAL_IS_F32 1
namespace MAVI_NS {
typedef float real;
typedef float2 real2;
__host__ __device__ __forceinline__ real2 make_real2(real x, real y) { return make_float2(x, y); }
// round-to-nearest / round-up products and sums that the compiler must not contract into FMAs
__device__ __forceinline__ real mul_ru(real a, real b) { return __fmul_ru(a, b); }
__device__ __forceinline__ real mul_rn(real a, real b) { return __fmul_rn(a, b); }
__device__ __forceinline__ real add_rn(real a, real b) { return __fadd_rn(a, b); }
// 1/x for a finite positive x (MUFU.RCP, 1 ulp)
__device__ __forceinline__ real fast_rcp(real x) {
  real r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
}  // namespace MAVI_NS
#else
#define MAVI_NS mavi_f64
#define MAVI_REAL_IS_F32 0
namespace MAVI_NS {
typedef double real;
typedef double2 real2;
__host__ __device__ __forceinline__ real2 make_real2(real x, real y) { return make_double2(x, y); }
__device__ __forceinline__ real mul_ru(real a, real b) { return __dmul_ru(a, b); }
__device__ __forceinline__ real mul_rn(real a, real b) { return __dmul_rn(a, b); }
__device__ __forceinline__ real add_rn(real a, real b) { return __dadd_rn(a, b); }
// 1/x to ~1 ulp without the special-case slow path of the IEEE division (x is a finite positive r^2 here).
__device__ __forceinline__ real fast_rcp(real x) {
  real r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));  // MUFU.RCP64H, ~20 bits
  real e = fma(-x, r, 1.0);   // |e| ~ 2^-20
  real t = fma(e, e, e);      // r (1 + e + e^2) = (1/x)(1 - e^3): one cubic step instead of two Newton steps
  return fma(r, t, r);
}
}  // namespace MAVI_NS
#endif
