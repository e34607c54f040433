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
// EPI = false: instead of L^-1, W receives the raw eliminated rows (W[k][r] = v_r[k]) and dout[r] = d_r; the caller finalises
// them with all its warps (potrf32_finalize) — the in-warp epilogue is 32 dependent shared-memory round trips per lane.
template <bool ROT, bool EPI = true>
__device__ __forceinline__ void potrf32_sym_t(const double* G, double* W, int* fail, double* dout = nullptr) {
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
  if (!EPI) {
    dout[r] = rsqrt_pivot(dr > 0.0 ? dr : 1.0);   // 1 / sqrt(d_r)
#pragma unroll
    for (int k = 0; k < HB; ++k) W[k * LDB + r] = v[k];
    return;
  }
  const double si = rsqrt_pivot(dr > 0.0 ? dr : 1.0);
  ssi[r] = si;
  __syncwarp();
  const double nrr = -(si * si);   // -1 / d_r
  // branch-free: the loads are unconditional (volatile), the triangle is cut by selects — a divergent branch per element
  // costs ~125 cycles here (measured: 4000 of 6900 cycles of the blocked variant were this loop)
  const volatile double* vssi = ssi;
#pragma unroll
  for (int k = 0; k < HB; ++k) {
    const double val = vssi[k] * (nrr * v[k]);
    const double w = (k > r) ? val : ((k == r) ? si : 0.0);
    W[k * LDB + r] = w;
  }
}

__device__ __forceinline__ void potrf32_sym(const double* G, double* W, int* fail) { potrf32_sym_t<true>(G, W, fail); }

// W (raw eliminated rows from potrf32_sym_t<.., false>: W[k][r] = v_r[k], sinv[r] = 1 / sqrt(d_r)) -> L^-1 in place, by nthreads
// threads (thread tid of them):  L^-1[k][r] = -v_r[k] / (sqrt(d_k) d_r) for k > r,  1 / sqrt(d_r) on the diagonal, 0 above it.
__device__ __forceinline__ void potrf32_finalize(double* W, const double* sinv, int tid, int nthreads) {
  for (int e = tid; e < HB * HB; e += nthreads) {
    const int k = e >> 5, r = e & 31;
    const double sik = sinv[k], sir = sinv[r];
    const double val = -(sik * (sir * sir)) * W[k * LDB + r];
    W[k * LDB + r] = (k > r) ? val : ((k == r) ? sik : 0.0);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Blocked variant (what the fused solve uses): the same symmetric elimination, in panels of 8 columns. Inside a panel every lane
// holds only its 8 panel entries in registers and a step updates at most 7 of them, so the warp issues almost nothing but the
// loop-carried chain (~100 cycles per column instead of ~190 with 30 FMAs per lane in its shadow); what the panel does to the
// columns right of it is ONE rank-8 update V[:, far] -= S C^T on the FP64 tensor pipe (mma.sync.m8n8k4.f64, 2 k-steps), with
//   S[r][j] = multiplier of row r at step j (0 for the pivot row),   C[k][j] = lane k's entry of panel column j before its step,
// both recorded in shared memory while the panel runs. V (32x32, stride LDB, both triangles valid) is updated in place and ends
// up holding, in row r, the unnormalised factor entries (k < r), d_r (k = r) and -d_r times column r of the inverse of the
// unit factor (k > r); W receives L^-1 as in potrf32_sym. One warp; no other warp is involved.
// ---------------------------------------------------------------------------------------------------------------------
template <int LDP, bool FAR, bool EPI, bool INPANEL>
__device__ __forceinline__ void potrf32_blk_t(double* V, double* W, int* fail) {
  __shared__ __align__(16) double Ct[HB * LDP];
  __shared__ __align__(16) double Sm[HB * LDP];
  __shared__ double ssi[HB];
  const int r = threadIdx.x & 31, g = r >> 2, tg = r & 3;
  const unsigned full = 0xffffffffu;
  __syncwarp();
  bool bad = false;
  double dr = 1.0;
#pragma unroll 1
  for (int c0 = 0; c0 < HB; c0 += 8) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) { const double2 t = *reinterpret_cast<const double2*>(V + r * LDB + c0 + j); v[j] = t.x; v[j + 1] = t.y; }
    double d = __shfl_sync(full, v[0], c0);
    double b = __shfl_sync(full, v[0], (c0 + 1) & 31);
    Ct[r * LDP] = v[0];
    __syncwarp();
    double colA[8], colB[8];   // column j for the updates of step j: loaded one step ahead (off the chain)
#pragma unroll
    for (int k = 2; k < 8; ++k) colA[k] = Ct[(c0 + k) * LDP];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      double* cur = (j & 1) ? colB : colA;
      double* nxt = (j & 1) ? colA : colB;
      bad |= !(d > 0.0);
      if (r == c) dr = d;
      const double y0 = rcp_seed(d);
      const double e = fma(-d, y0, 1.0);
      const double vc = (r == c) ? 0.0 : v[j];     // the pivot row itself is left alone
      const double t = fma(e, e, e);
      double dn = 0.0, bn = 0.0;
      if (j < 7) {
        const double p1 = vc * b, q0 = p1 * y0, a0 = fma(-p1, y0, v[j + 1]);
        v[j + 1] = fma(-q0, t, a0);                // = v - vc b / d_c: the next pivot's row entry, 3 FMAs behind the seed
        dn = __shfl_sync(full, v[j + 1], (c + 1) & 31);
        bn = __shfl_sync(full, v[j + 1], (c + 2) & 31);
        Ct[r * LDP + j + 1] = v[j + 1];            // column j+1 is final: publish it for the next step and for the rank-8 update
      }
      const double s = vc * fma(t, y0, y0);
      Sm[r * LDP + j] = s;
      __syncwarp();
      if (INPANEL) {
      if (j < 6) {
#pragma unroll
        for (int k = j + 3; k < 8; ++k) nxt[k] = Ct[(c0 + k) * LDP + j + 1];
      }
#pragma unroll
      for (int k = j + 2; k < 8; ++k) v[k] = fma(-s, cur[k], v[k]);
      }
      d = dn; b = bn;
    }
    // the panel's own columns are final
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(V + r * LDB + c0 + j) = make_double2(v[j], v[j + 1]);
    // rank-8 update of the columns right of the panel: per column tile 4 row tiles = 4 independent chains of 2 MMAs
    if (FAR) {
    double fa[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) { fa[m][0] = Sm[(8 * m + g) * LDP + tg]; fa[m][1] = Sm[(8 * m + g) * LDP + 4 + tg]; }
#pragma unroll 1
    for (int n0 = c0 + 8; n0 < HB; n0 += 8) {
      const double fb0 = Ct[(n0 + g) * LDP + tg], fb1 = Ct[(n0 + g) * LDP + 4 + tg];
      double2 cv[4];
      double x[4][2];
#pragma unroll
      for (int m = 0; m < 4; ++m) { cv[m] = *reinterpret_cast<const double2*>(V + (8 * m + g) * LDB + n0 + 2 * tg); x[m][0] = x[m][1] = 0.0; }
#pragma unroll
      for (int m = 0; m < 4; ++m) dmma_m8n8k4(x[m][0], x[m][1], fa[m][0], fb0);
#pragma unroll
      for (int m = 0; m < 4; ++m) dmma_m8n8k4(x[m][0], x[m][1], fa[m][1], fb1);
#pragma unroll
      for (int m = 0; m < 4; ++m) *reinterpret_cast<double2*>(V + (8 * m + g) * LDB + n0 + 2 * tg) = make_double2(cv[m].x - x[m][0], cv[m].y - x[m][1]);
    }
    }
    __syncwarp();
  }
  if (bad && r == 0) atomicExch(fail, 1);
  const double si = rsqrt_pivot(dr > 0.0 ? dr : 1.0);
  ssi[r] = si;
  __syncwarp();
  const double nrr = -(si * si);   // -1 / d_r
  if (EPI) {
    const volatile double* vssi = ssi;
    const volatile double* vV = V;
#pragma unroll 8
    for (int it = 0; it < HB; ++it) {   // skewed so that neither the row read of V nor the column write of W has bank conflicts; branch-free
      const int k = (r + it) & 31;
      const double val = vssi[k] * (nrr * vV[r * LDB + k]);
      const double w = (k > r) ? val : ((k == r) ? si : 0.0);
      W[k * LDB + r] = w;
    }
  }
}



__device__ __forceinline__ void potrf32_blk(double* V, double* W, int* fail) { potrf32_blk_t<12, true, true, true>(V, W, fail); }

// ---------------------------------------------------------------------------------------------------------------------
// Two-warp form of the blocked variant: the chain warp (role 0) only applies a panel's rank-8 update to the NEXT panel's eight
// columns (4 tiles); a helper warp (role 1) applies it to the columns beyond while the chain warp is already in the next
// panel. The panel records are double buffered by panel parity; two named barriers (the chain warp only ever arrives on the
// first and only waits on the second when it needs columns the helper has written) carry the hand-offs. Both warps call this
// with the same arguments; the warps of the CTA that do not take part must not use barriers bar0, bar0 + 1.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void potrf32_blk2(double* V, double* W, int* fail, int role, int bar0) {
  constexpr int LDP = 12;
  __shared__ __align__(16) double Ct[2][HB * LDP];
  __shared__ __align__(16) double Sm[2][HB * LDP];
  __shared__ double ssi[HB];
  const int r = threadIdx.x & 31, g = r >> 2, tg = r & 3;
  const unsigned full = 0xffffffffu;
  __syncwarp();
  // rank-8 update of column tile n0 with the records of panel parity pp (4 independent chains of 2 MMAs)
  auto far_tile = [&](int pp, int n0) {
    const double* C = Ct[pp];
    const double* S = Sm[pp];
    const double fb0 = C[(n0 + g) * LDP + tg], fb1 = C[(n0 + g) * LDP + 4 + tg];
    double2 cv[4];
    double x[4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m) { cv[m] = *reinterpret_cast<const double2*>(V + (8 * m + g) * LDB + n0 + 2 * tg); x[m][0] = x[m][1] = 0.0; }
#pragma unroll
    for (int m = 0; m < 4; ++m) dmma_m8n8k4(x[m][0], x[m][1], S[(8 * m + g) * LDP + tg], fb0);
#pragma unroll
    for (int m = 0; m < 4; ++m) dmma_m8n8k4(x[m][0], x[m][1], S[(8 * m + g) * LDP + 4 + tg], fb1);
#pragma unroll
    for (int m = 0; m < 4; ++m) *reinterpret_cast<double2*>(V + (8 * m + g) * LDB + n0 + 2 * tg) = make_double2(cv[m].x - x[m][0], cv[m].y - x[m][1]);
  };
  if (role == 1) {
    // helper: panels 0 and 1 have columns beyond the next panel (16.. and 24..)
    for (int p = 0; p < 2; ++p) {
      bar_sync_n(bar0, 64);                       // records of panel p are complete
      for (int n0 = 8 * p + 16; n0 < HB; n0 += 8) far_tile(p & 1, n0);
      __syncwarp();
      bar_arrive_n(bar0 + 1, 64);                 // columns >= 8 p + 16 carry panel p
    }
    return;
  }
  bool bad = false;
  double dr = 1.0;
#pragma unroll 1
  for (int c0 = 0; c0 < HB; c0 += 8) {
    const int pp = (c0 >> 3) & 1;
    double* C = Ct[pp];
    double* S = Sm[pp];
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) { const double2 t = *reinterpret_cast<const double2*>(V + r * LDB + c0 + j); v[j] = t.x; v[j + 1] = t.y; }
    double d = __shfl_sync(full, v[0], c0);
    double b = __shfl_sync(full, v[0], (c0 + 1) & 31);
    C[r * LDP] = v[0];
    __syncwarp();
    double colA[8], colB[8];   // column j for the updates of step j: loaded one step ahead (off the chain)
#pragma unroll
    for (int k = 2; k < 8; ++k) colA[k] = C[(c0 + k) * LDP];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      double* cur = (j & 1) ? colB : colA;
      double* nxt = (j & 1) ? colA : colB;
      bad |= !(d > 0.0);
      if (r == c) dr = d;
      const double y0 = rcp_seed(d);
      const double e = fma(-d, y0, 1.0);
      const double vc = (r == c) ? 0.0 : v[j];     // the pivot row itself is left alone
      const double t = fma(e, e, e);
      double dn = 0.0, bn = 0.0;
      if (j < 7) {
        const double p1 = vc * b, q0 = p1 * y0, a0 = fma(-p1, y0, v[j + 1]);
        v[j + 1] = fma(-q0, t, a0);                // = v - vc b / d_c: the next pivot's row entry, 3 FMAs behind the seed
        dn = __shfl_sync(full, v[j + 1], (c + 1) & 31);
        bn = __shfl_sync(full, v[j + 1], (c + 2) & 31);
        C[r * LDP + j + 1] = v[j + 1];             // column j+1 is final: publish it for the next step and for the rank-8 update
      }
      const double s = vc * fma(t, y0, y0);
      S[r * LDP + j] = s;
      __syncwarp();
      if (j < 6) {
#pragma unroll
        for (int k = j + 3; k < 8; ++k) nxt[k] = C[(c0 + k) * LDP + j + 1];
      }
#pragma unroll
      for (int k = j + 2; k < 8; ++k) v[k] = fma(-s, cur[k], v[k]);
      d = dn; b = bn;
    }
    // the panel's own columns are final
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(V + r * LDB + c0 + j) = make_double2(v[j], v[j + 1]);
    __syncwarp();
    if (c0 < 16) bar_arrive_n(bar0, 64);          // the helper may take the columns beyond the next panel
    // the helper's update with the PREVIOUS panel touches the next panel's columns too: it must have landed before they are
    // read-modify-written here (it had this whole panel's time)
    if (c0 >= 8 && c0 + 8 < HB) bar_sync_n(bar0 + 1, 64);
    if (c0 + 8 < HB) far_tile(pp, c0 + 8);        // the next panel's columns: on the chain
    __syncwarp();
  }
  if (bad && r == 0) atomicExch(fail, 1);
  const double si = rsqrt_pivot(dr > 0.0 ? dr : 1.0);
  ssi[r] = si;
  __syncwarp();
  const double nrr = -(si * si);   // -1 / d_r
  const volatile double* vssi = ssi;
  const volatile double* vV = V;
#pragma unroll 8
  for (int it = 0; it < HB; ++it) {   // skewed so that neither the row read of V nor the column write of W has bank conflicts; branch-free
    const int k = (r + it) & 31;
    const double val = vssi[k] * (nrr * vV[r * LDB + k]);
    const double w = (k > r) ? val : ((k == r) ? si : 0.0);
    W[k * LDB + r] = w;
  }
}



}  // namespace tsl
