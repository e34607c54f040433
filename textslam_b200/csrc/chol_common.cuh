// Constants and the two primitives shared by the wave kernels (chol.cu, chol_tile.cuh) and the fused solve (chol_fused.cu).
#pragma once
#include <cuda_runtime.h>

namespace tsl {

constexpr int NB = 64;        // tile of the reduced camera matrix
constexpr int SPAD = NB + 4;  // smem row stride (doubles): conflict-free m8n8k4 fragment loads
constexpr int HB = 32;        // half block

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// 1/sqrt(d), branch-free: MUFU.RSQ64H seed (PTX rsqrt.approx.ftz.f64, ~2^-22) and one cubic (Householder) step
// y = y0 + y0 e (1/2 + 3/8 e), e = 1 - d y0^2  -> relative error ~ e^3, i.e. rounding level.
__device__ __forceinline__ double rsqrt_pivot(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-d, y0 * y0, 1.0);
  return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}

}  // namespace tsl
