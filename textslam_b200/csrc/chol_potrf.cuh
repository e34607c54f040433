// The 32x32 diagonal-block routine of the fused reduced-system solve (chol_fused.cu) — in a header so that
// tools/ubench/potrf_sym_bench.cu can time and verify it in isolation.
#pragma once
#include "chol_common.cuh"

namespace tsl {

constexpr int LDB = 36;              // shared-memory stride of a 32x32 block (doubles): = 4 mod 16 -> conflict-free m8n8k4 fragment loads

// 1/d to rounding level, branch-free: MUFU.RCP64H seed (PTX rcp.approx.ftz.f64, ~2^-20) + one cubic step y0 (1 + e + e^2)
__device__ __forceinline__ double rcp_seed(double d) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  return y0;
}
__device__ __forceinline__ double rcp_pivot(double d) {
  const double y0 = rcp_seed(d);
  const double e = fma(-d, y0, 1.0);
  return fma(fma(e, e, e), y0, y0);
}

// ---------------------------------------------------------------------------------------------------------------------
// 32x32 diagonal block by ONE warp. G: the block in shared memory (stride LDB), SYMMETRIC (both triangles valid).
// Lane r holds the full row r of the symmetric matrix and every step c applies the rank-1 elimination update to ALL rows but
// the pivot row, for the columns k > c:   v_r[k] -= (v_r[c] / d_c) v_k[c].   Rows r > c carry the Schur complement (their
// entries k <= r are the unnormalised factor columns), rows r < c carry -d_r times column r of the inverse of the unit factor —
// the recurrence of the forward substitution M N = I is the same update — so at the end
//   L[r][k] = v_r[k] / sqrt(d_k) (k < r),     L^-1[k][r] = -v_r[k] / (sqrt(d_k) d_r) (k > r),    L^-1[r][r] = 1 / sqrt(d_r).
// Only the chain  d_c -> 1/d_c -> one FMA on lane c+1 -> shuffle  is loop carried; the column broadcast (shared memory, double
// buffered) and the 31-c FMAs per lane fill its stall slots. W receives L^-1 as a full row-major block (zeros above the diagonal).
// A non-positive pivot raises *fail and is replaced by 1 (the caller rejects the step).
// ---------------------------------------------------------------------------------------------------------------------
template <bool ROT>
__device__ __forceinline__ void potrf32_sym_t(const double* G, double* W, int* fail) {
  // Column buffers are twice as long as a column (entry k also lives at k + 32): the loop below works in a register frame that
  // is rotated by 8 columns per trip, and base + position indexes the doubled buffer without a wrap.
  __shared__ __align__(16) double colbuf[2][2 * HB];
  __shared__ double ssi[HB];
  const int r = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  __syncwarp();
  double v[HB];
#pragma unroll
  for (int k = 0; k < HB; ++k) v[k] = G[k * LDB + r];   // column r = row r
  // Software pipeline, one region per eliminated column between two warp barriers: the region of column c applies column c with
  // 1/d_c from the region before, and computes d_{c+1}, its reciprocal and the two broadcast scalars of the next region while the
  // update FMAs of every lane fill the stall slots of that chain.
  //   d  = pivot d_c (lane c's diagonal),  b = lane (c+1)'s entry of column c  (both by shuffle: they sit on the chain)
  //   colbuf[c & 1][k] = lane k's entry of column c, ZERO for k < c (shared memory: feeds the FMAs, off the chain)
  // Code size matters more than FMA count here: 32 fully unrolled columns are ~28 KB of straight-line SASS that one warp
  // streams through once per call (the instruction fetch then sets the pace: measured 260 cycles per column). Instead the
  // columns go in 4 trips of 8 through ONE unrolled body of 8 columns (~10 KB, resident), with the register frame rotated by
  // 8 positions after each trip: the pivot of column 8 it + j always sits at position j. Positions that have wrapped around
  // hold finished columns; their update is a multiplication by the zero the finished lanes publish.
  bool bad = false;
  double d = __shfl_sync(full, v[0], 0);
  double b = __shfl_sync(full, v[0], 1);
  colbuf[0][r] = v[0]; colbuf[0][r + HB] = v[0];
  bad |= !(d > 0.0);
  d = d > 0.0 ? d : 1.0;
  double dr = d;                               // lane 0 keeps d_0; the others overwrite it at their own pivot
  double rinv = rcp_pivot(d);
  double vc = (r == 0) ? 0.0 : v[0];           // the pivot row itself is left alone
  double p1 = vc * b;
  __syncwarp();
  if (ROT) {
  // The chain per column is  MUFU seed y0 -> e = 1 - d y0 -> t = e + e^2 -> x = (v - p1 y0) - (p1 y0) t -> shuffle:  the last FMA
  // of the reciprocal is folded into the update of lane c+1's next pivot (1/d = y0 + y0 t), p1 y0 and v - p1 y0 are formed beside
  // e and t. The full reciprocal (for the other columns' multiplier s) is off the chain, and so is the positivity test: a
  // non-positive pivot only raises `bad`, the numbers that follow it are garbage either way and the caller rejects the step.
  double colA[HB], colB[HB];
#pragma unroll
  for (int p = 2; p < HB; p += 2) { const double2 t = *reinterpret_cast<const double2*>(&colbuf[0][p]); colA[p] = t.x; colA[p + 1] = t.y; }
  double y0 = rcp_seed(d);
#pragma unroll 1
  for (int base = 0; base < HB; base += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = base + j;
      // (column 31 has nothing left to update: its region runs with vc = 0, i.e. as a no-op, to keep the body branch-free)
      double* cur = (j & 1) ? colB : colA;     // column c, loaded during the region before
      double* nxt = (j & 1) ? colA : colB;
      const double e = fma(-d, y0, 1.0);
      const double q0 = p1 * y0;
      const double a0 = fma(-p1, y0, v[j + 1]);
      const double t = fma(e, e, e);
      v[j + 1] = fma(-q0, t, a0);                // = v - p1 / d_c
      const double dn = __shfl_sync(full, v[j + 1], (c + 1) & 31);
      const double y0n = rcp_seed(dn);
      const double bn = __shfl_sync(full, v[j + 1], (c + 2) & 31);
      const double rinv = fma(t, y0, y0);
      const double s = vc * rinv;
      const double pv = (r > c) ? v[j + 1] : 0.0;
      colbuf[(c + 1) & 1][r] = pv; colbuf[(c + 1) & 1][r + HB] = pv;
      __syncwarp();
      {   // column c+1 for the next region, in flight while this region's FMAs issue (frame of the next region: rotated after j = 7)
        const int jn = (j + 1) & 7;
        const double* cb = &colbuf[(c + 1) & 1][c + 1 - jn];
#pragma unroll
        for (int p = (jn + 2) & ~1; p < HB; p += 2) { const double2 t2 = *reinterpret_cast<const double2*>(cb + p); nxt[p] = t2.x; nxt[p + 1] = t2.y; }
      }
      bad |= (c + 1 < HB) & !(dn > 0.0);
      if (r == c + 1) dr = dn;
#pragma unroll
      for (int p = j + 2; p < HB; ++p) v[p] = fma(-s, cur[p], v[p]);
      vc = (r == c + 1 || c + 2 >= HB) ? 0.0 : v[j + 1];
      p1 = vc * bn;
      d = dn; y0 = y0n;
    }
    // rotate the frame by 8 positions
    double t8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) t8[q] = v[q];
#pragma unroll
    for (int p = 0; p < HB - 8; ++p) v[p] = v[p + 8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[HB - 8 + q] = t8[q];
  }
  } else {
    // straight-line variant: 31 regions, region c touches only the columns k > c + 1 (half the FMAs of the rotating frame,
    // three times its code size)
#pragma unroll
    for (int c = 0; c + 1 < HB; ++c) {
      double col[HB];
#pragma unroll
      for (int k = (c + 2) & ~1; k < HB; k += 2) { const double2 t = *reinterpret_cast<const double2*>(&colbuf[c & 1][k]); col[k] = t.x; col[k + 1] = t.y; }
      v[c + 1] = fma(-p1, rinv, v[c + 1]);
      const double s = vc * rinv;
      double dn = __shfl_sync(full, v[c + 1], c + 1);
      const double bn = (c + 2 < HB) ? __shfl_sync(full, v[c + 1], c + 2) : 0.0;
      colbuf[(c + 1) & 1][r] = v[c + 1];
      bad |= !(dn > 0.0);
      dn = dn > 0.0 ? dn : 1.0;
      if (r == c + 1) dr = dn;
      const double rinv_n = rcp_pivot(dn);
#pragma unroll
      for (int k = c + 2; k < HB; ++k) v[k] = fma(-s, col[k], v[k]);
      vc = (r == c + 1) ? 0.0 : v[c + 1];
      p1 = vc * bn;
      rinv = rinv_n;
      __syncwarp();
    }
  }
  if (bad && r == 0) atomicExch(fail, 1);
  const double si = rsqrt_pivot(dr > 0.0 ? dr : 1.0);
  ssi[r] = si;
  __syncwarp();
  const double nrr = -(si * si);   // -1 / d_r
#pragma unroll
  for (int k = 0; k < HB; ++k) {
    const double w = (k < r) ? 0.0 : ((k == r) ? si : ssi[k] * (nrr * v[k]));
    W[k * LDB + r] = w;
  }
}

__device__ __forceinline__ void potrf32_sym(const double* G, double* W, int* fail) { potrf32_sym_t<true>(G, W, fail); }

}  // namespace tsl
