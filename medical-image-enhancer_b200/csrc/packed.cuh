// Packed float32x2 arithmetic of sm_100a (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2):
// one issue slot, two independently IEEE-rounded results, operand negation folded in.  A float2
// holds the two lanes (x = first pixel).  Used by the kernels that process two pixels per thread.
#pragma once
#include "common.cuh"

namespace mdimg {

#ifdef __CUDACC__
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float2 a) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return upk(r); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return upk(r); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return upk(r); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return upk(r); }
// ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false; where the
// reference rounds the product and the sum separately the sum is issued as two scalar add.rn.f32,
// which are never contracted.
__device__ __forceinline__ float2 add2_nc(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }     // folds into the consumer
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float rsq_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif

}  // namespace mdimg
