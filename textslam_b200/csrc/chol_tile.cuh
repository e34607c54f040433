// Device tile routines of the reduced-system Cholesky (chol.cu) — kept in a header so that tools/ubench/chol_tile_bench.cu
// can time and verify them in isolation.
#pragma once
#include <cuda_runtime.h>
#include "chol_common.cuh"

namespace tsl {


// ---------------------------------------------------------------------------------------------
// device tile routines for the 64x64 diagonal tile (CTA = 128 threads, tile and right-hand side in shared memory).
// The factorisation is recursive over 32x32 blocks: the sequential part (32-column Crout / substitution with the
// row held in registers, fully unrolled) exists ONCE as a __noinline__ function and is called twice, the coupling
// between the halves is a small register-tiled GEMM. A flat 64-column unrolled version measured 64 us per launch
// because 3 x 2016 FMAs of straight-line code miss the instruction cache (profiles/r1_notes.md).
// ---------------------------------------------------------------------------------------------
constexpr int LDT = NB + 1;        // smem leading dimension (doubles)
constexpr int PT_THREADS = 128;    // CTA size of potrf_trsm_kernel

// Crout Cholesky of the 32x32 block at M (lower, in place). Threads 0..31 own one row each (registers); finished
// entries are published to M so that row c is read as a broadcast. sinv[c] = 1 / L[c][c]. All CTA threads call it.
__device__ __noinline__ void potrf32(double* M, double* sinv, int* fail) {
  const int r = threadIdx.x;
  const bool owner = r < HB;
  double row[HB];
#pragma unroll
  for (int c = 0; c < HB; ++c) row[c] = (owner && c <= r) ? M[r * LDT + c] : 0.0;
#pragma unroll
  for (int c = 0; c < HB; ++c) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (owner && r >= c) {
#pragma unroll
      for (int k = 0; k + 3 < c; k += 4) {
        s0 += row[k] * M[c * LDT + k]; s1 += row[k + 1] * M[c * LDT + k + 1];
        s2 += row[k + 2] * M[c * LDT + k + 2]; s3 += row[k + 3] * M[c * LDT + k + 3];
      }
#pragma unroll
      for (int k = c & ~3; k < c; ++k) s0 += row[k] * M[c * LDT + k];
    }
    const double s = row[c] - ((s0 + s1) + (s2 + s3));
    if (r == c) {
      if (!(s > 0.0)) atomicExch(fail, 1);  // not positive definite (or NaN): report, continue with a harmless pivot
      sinv[c] = (s > 0.0) ? rsqrt(s) : 1.0;
    }
    __syncthreads();
    if (owner && r >= c) { row[c] = s * sinv[c]; M[r * LDT + c] = row[c]; }   // diagonal: s * rsqrt(s) = sqrt(s)
    __syncthreads();
  }
}

// X L^T = B for `nrows` (<= 64) rows and a 32x32 lower block L (both in shared memory, in place on X).
// Thread r owns row r in registers; rows are independent, no barrier inside.
__device__ __noinline__ void trsm32(double* X, int nrows, const double* L, const double* sinv) {
  const int r = threadIdx.x;
  if (r >= nrows) return;
  double x[HB];
#pragma unroll
  for (int c = 0; c < HB; ++c) x[c] = X[r * LDT + c];
#pragma unroll
  for (int c = 0; c < HB; ++c) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k + 3 < c; k += 4) {
      s0 += x[k] * L[c * LDT + k]; s1 += x[k + 1] * L[c * LDT + k + 1];
      s2 += x[k + 2] * L[c * LDT + k + 2]; s3 += x[k + 3] * L[c * LDT + k + 3];
    }
#pragma unroll
    for (int k = c & ~3; k < c; ++k) s0 += x[k] * L[c * LDT + k];
    x[c] = (x[c] - ((s0 + s1) + (s2 + s3))) * sinv[c];
  }
#pragma unroll
  for (int c = 0; c < HB; ++c) X[r * LDT + c] = x[c];
}

// C[m x 32] -= A[m x 32] B[32 x 32]^T, everything in shared memory (ld LDT), 4x4 register tiles, m in {32, 64}.
__device__ __noinline__ void gemm_nt32(double* C, const double* A, const double* B, int m) {
  const int nb = (m / 4) * (HB / 4);
  for (int blk = threadIdx.x; blk < nb; blk += PT_THREADS) {
    const int bi = blk / (HB / 4), bj = blk - bi * (HB / 4);
    const double* a = A + 4 * bi * LDT; const double* b = B + 4 * bj * LDT;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
#pragma unroll 4
    for (int k = 0; k < HB; ++k) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = a[i * LDT + k]; bv[i] = b[i * LDT + k]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] += av[i] * bv[jj];
    }
    double cv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) cv[i][jj] = C[(4 * bi + i) * LDT + 4 * bj + jj];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) C[(4 * bi + i) * LDT + 4 * bj + jj] = cv[i][jj] - acc[i][jj];
  }
}

// 64x64 Cholesky in place on sT (lower):  L11 = chol(A11); L21 = A21 L11^-T; A22 -= L21 L21^T; L22 = chol(A22)
__device__ __forceinline__ void potrf_tile(double* sT, double* sinv, int* fail) {
  potrf32(sT, sinv, fail);
  trsm32(sT + HB * LDT, HB, sT, sinv);
  __syncthreads();
  gemm_nt32(sT + HB * LDT + HB, sT + HB * LDT, sT + HB * LDT, HB);
  __syncthreads();
  potrf32(sT + HB * LDT + HB, sinv + HB, fail);
}

// X L^T = B for a 64-row tile sX (in place) against the factored sT:
//   X1 = B1 L11^-T ;  B2 -= X1 L21^T ;  X2 = B2 L22^-T
__device__ __forceinline__ void trsm_tile(double* sX, const double* sT, const double* sinv) {
  trsm32(sX, NB, sT, sinv);
  __syncthreads();
  gemm_nt32(sX + HB, sX, sT + HB * LDT, NB);
  __syncthreads();
  trsm32(sX + HB, NB, sT + HB * LDT + HB, sinv + HB);
  __syncthreads();
}


// C (64x64 at C, ld) -= Xi Xk^T with Xi, Xk 64x64 tiles (ld). 4 warps (2x2), warp tile 32x32.
__device__ __forceinline__ void gemm_tile_nt(const double* __restrict__ Xi, const double* __restrict__ Xk, double* __restrict__ C, int ld,
                                             double* sA, double* sB) {
  // 2 x 32 KB tile loads: all 16-byte loads of a batch are issued before the first shared-memory store
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    double2 va[8], vb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = threadIdx.x + 128 * (8 * half + u), r = e >> 5, c2 = (e & 31) * 2;
      va[u] = *reinterpret_cast<const double2*>(Xi + (size_t)r * ld + c2);
      vb[u] = *reinterpret_cast<const double2*>(Xk + (size_t)r * ld + c2);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = threadIdx.x + 128 * (8 * half + u), r = e >> 5, c2 = (e & 31) * 2;
      sA[r * SPAD + c2] = va[u].x; sA[r * SPAD + c2 + 1] = va[u].y;
      sB[r * SPAD + c2] = vb[u].x; sB[r * SPAD + c2 + 1] = vb[u].y;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
  const int g = lane >> 2, tg = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < NB; k0 += 4) {
    double fa[4], fb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) fa[a] = sA[(wr + 8 * a + g) * SPAD + k0 + tg];
#pragma unroll
    for (int b = 0; b < 4; ++b) fb[b] = sB[(wc + 8 * b + g) * SPAD + k0 + tg];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = wr + 8 * a + g, c = wc + 8 * b + 2 * tg;
      double2* p = reinterpret_cast<double2*>(C + (size_t)r * ld + c);
      double2 v = *p;
      v.x -= acc[a][b][0]; v.y -= acc[a][b][1];
      *p = v;
    }
}


// =============================================================================================
// Second-generation tile routines (measured with tools/ubench/chol_tile_bench.cu, profiles/r1_notes.md).
// What the first generation lost its time on: every column of the Crout elimination re-read a whole row of the
// factor from shared memory just in time (LDS -> DFMA serialised, ~430 cycles per column, 2 x 13.9k cycles per 64-tile),
// with two block barriers per column. Here the 32x32 factorisation is right-looking inside ONE warp, branch-free:
// the dependent chain of a column is rsqrt -> scale -> own next diagonal -> shuffle broadcast, everything else
// (publishing the column, the rank-1 update with batched 128-bit broadcast loads) fills its stall slots.
// Shared-memory tiles use an EVEN leading dimension LD2 (16-byte aligned pairs -> LDS.128).
// =============================================================================================
constexpr int LD2 = NB + 2;


// Cholesky of the 32x32 block at M (lower, in place, leading dimension LD2) by ONE warp; lane r owns row r in
// registers. Also writes the transpose of the factor to Mt (Mt[c][r] = L[r][c]) for the right-looking solves and
// sinv[c] = 1 / L[c][c]. A non-positive pivot raises *fail and is replaced by 1 (the caller rejects the step).
__device__ __noinline__ void potrf32_rl(double* M, double* Mt, double* sinv, int* fail) {
  __shared__ __align__(16) double colbuf[2][HB];
  const int r = threadIdx.x & 31;
  __syncwarp();   // the shuffles below need the whole warp converged (a lane-divergent caller would take the slow path)
  double row[HB];
#pragma unroll
  for (int c = 0; c < HB; ++c) row[c] = (c <= r) ? M[r * LD2 + c] : 0.0;
  double d = __shfl_sync(0xffffffffu, row[0], 0);
  bool bad = false;
#pragma unroll
  for (int c = 0; c < HB; ++c) {
    const bool ok = d > 0.0;
    bad |= !ok;
    const double si = rsqrt_pivot(ok ? d : 1.0);
    const double l = row[c] * si;            // lanes r < c carry don't-care values in row[c..]; they are never stored
    row[c] = l;
    if (r == c) sinv[c] = si;
    if (c + 1 < HB) {
      d = __shfl_sync(0xffffffffu, fma(-l, l, row[c + 1]), c + 1);   // next pivot: lane c+1 needs only its own l
      colbuf[c & 1][r] = l;
      __syncwarp();
      double lk[HB];
#pragma unroll
      for (int k = (c + 1) & ~1; k < HB; k += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&colbuf[c & 1][k]);
        lk[k] = v.x; lk[k + 1] = v.y;
      }
#pragma unroll
      for (int k = c + 1; k < HB; ++k) row[k] = fma(-l, lk[k], row[k]);
    }
  }
  if (bad && r == 0) atomicExch(fail, 1);
#pragma unroll
  for (int c = 0; c < HB; ++c) if (c <= r) { M[r * LD2 + c] = row[c]; Mt[c * LD2 + r] = row[c]; }
}

// X L^T = B for one row (in place on X[0..32)) against a 32x32 lower factor given as its TRANSPOSE Lt (Lt[c][k] =
// L[k][c], leading dimension LD2): right-looking, once x[c] is final it is subtracted from every later column
// (independent FMAs; the dependent chain per column is one multiply + one FMA).
__device__ __noinline__ void trsm32_row(double* X, const double* Lt, const double* sinv) {
  double x[HB];
#pragma unroll
  for (int c = 0; c < HB; c += 2) { const double2 v = *reinterpret_cast<const double2*>(X + c); x[c] = v.x; x[c + 1] = v.y; }
#pragma unroll
  for (int c = 0; c < HB; ++c) {
    x[c] *= sinv[c];
    double lk[HB];
#pragma unroll
    for (int k = (c + 1) & ~1; k < HB; k += 2) {
      const double2 v = *reinterpret_cast<const double2*>(Lt + c * LD2 + k);
      lk[k] = v.x; lk[k + 1] = v.y;
    }
#pragma unroll
    for (int k = c + 1; k < HB; ++k) x[k] = fma(-x[c], lk[k], x[k]);
  }
#pragma unroll
  for (int c = 0; c < HB; c += 2) *reinterpret_cast<double2*>(X + c) = make_double2(x[c], x[c + 1]);
}

// C[m x 32] -= A[m x 32] B[32 x 32]^T in shared memory (leading dimension LD2), m = 8 RI... precisely: the calling
// group has 8 * (m / RI) threads, tid in [0, 8 m / RI): thread (rg, cg) owns rows rg + (m/RI) i and columns cg + 8 j
// (interleaved so that the 8 lanes of a 128-bit load phase hit 8 consecutive rows = all 32 banks).
template <int RI>
__device__ __noinline__ void gemm_nt32_il(double* C, const double* A, const double* B, int m, int tid) {
  const int nrg = m / RI;
  const int cg = tid & 7, rg = tid >> 3;
  double acc[RI][4];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
  for (int k = 0; k < HB; k += 2) {
    double2 av[RI], bv[4];
#pragma unroll
    for (int i = 0; i < RI; ++i) av[i] = *reinterpret_cast<const double2*>(A + (rg + nrg * i) * LD2 + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const double2*>(B + (cg + 8 * j) * LD2 + k);
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i].y, bv[j].y, fma(av[i].x, bv[j].x, acc[i][j]));
  }
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) C[(rg + nrg * i) * LD2 + cg + 8 * j] -= acc[i][j];
}

// Phase 3 of factor_solve_tile on the FP64 tensor pipe: A22 (32 x 32) -= L21 L21^T and X2 (64 x 32) -= X1 L21^T as twelve
// 8-row strips (4 of A22, 8 of X2), three per warp; each strip is 4 column tiles x 8 k-steps of mma.m8n8k4 (fragment layout
// as in gemm_tile_nt). The register-tiled FMA version (gemm_nt32_il) spent 4.1k cycles here, bound by shared-memory bandwidth
// (one 128-bit load per 4 FMAs); the MMA fragments need one 64-bit load per 64 FMAs.
__device__ __forceinline__ void update_phase_dmma(double* sT, double* sX, int tid) {
  const int w = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const double* B = sT + HB * LD2;                       // L21: rows 32..63, columns 0..31
#pragma unroll 1
  for (int t = w; t < 12; t += 4) {
    const double* A = t < 4 ? sT + (HB + 8 * t) * LD2 : sX + 8 * (t - 4) * LD2;   // strip of L21 / of X1 (columns 0..31)
    double* C = (t < 4 ? sT + (HB + 8 * t) * LD2 : sX + 8 * (t - 4) * LD2) + HB; // same rows, columns 32..63
    double acc[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c][0] = acc[c][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < HB; k0 += 4) {
      const double fa = A[g * LD2 + k0 + tg];
#pragma unroll
      for (int c = 0; c < 4; ++c) dmma_m8n8k4(acc[c][0], acc[c][1], fa, B[(8 * c + g) * LD2 + k0 + tg]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double2* p = reinterpret_cast<double2*>(C + g * LD2 + 8 * c + 2 * tg);
      double2 v = *p;
      v.x -= acc[c][0]; v.y -= acc[c][1];
      *p = v;
    }
  }
}

// Factor the 64x64 diagonal tile in sT (lower, LD2) and solve X L^T = B for the 64 rows of sX (128-thread CTA):
//   phase 1  warp 0: L11 = chol(A11)
//   phase 2  warp 0: L21 = A21 L11^-T            | warps 2,3: X1 = B1 L11^-T        (same code on both sides)
//   phase 3  all   : A22 -= L21 L21^T ;  B2 -= X1 L21^T
//   phase 4  warp 0: L22 = chol(A22)
//   phase 5  warps 0,1: X2 = B2 L22^-T
// sLt receives L11^T and L22^T (diagonal blocks only). Measured (tools/ubench/chol_tile_bench.cu): running the B2 update
// on warps 2,3 WHILE warp 0 factors A22 is 3x slower than doing them one after the other (30.2k vs 7.2k + 2.5k cycles):
// two different unrolled instruction streams on one SM evict each other from the instruction cache, so phases only
// ever overlap identical code. STAMPS: optional clock64() trace for the micro-benchmark.
template <bool STAMPS, bool MMA_UPDATE = false>
__device__ __forceinline__ void factor_solve_tile(double* sT, double* sX, double* sLt, double* sinv, int* fail, long long* stamps) {
  const int tid = threadIdx.x, warp = tid >> 5;
#define TSL_STAMP(k) do { if (STAMPS && tid == 0) stamps[k] = clock64(); } while (0)
  if (warp == 0) potrf32_rl(sT, sLt, sinv, fail);
  __syncthreads();
  TSL_STAMP(2);
  if (warp == 0) trsm32_row(sT + (HB + tid) * LD2, sLt, sinv);
  else if (warp >= 2) trsm32_row(sX + (tid - 64) * LD2, sLt, sinv);
  __syncthreads();
  TSL_STAMP(3);
  if (MMA_UPDATE) update_phase_dmma(sT, sX, tid);
  else {
    gemm_nt32_il<2>(sT + HB * LD2 + HB, sT + HB * LD2, sT + HB * LD2, HB, tid);
    gemm_nt32_il<4>(sX + HB, sX, sT + HB * LD2, NB, tid);
  }
  __syncthreads();
  TSL_STAMP(4);
  if (warp == 0) potrf32_rl(sT + HB * LD2 + HB, sLt + HB * LD2 + HB, sinv + HB, fail);
  __syncthreads();
  TSL_STAMP(5);
  if (warp < 2) trsm32_row(sX + tid * LD2 + HB, sLt + HB * LD2 + HB, sinv + HB);
  __syncthreads();
  TSL_STAMP(6);
#undef TSL_STAMP
}

// ---------------------------------------------------------------------------------------------
// Two-tile panels (potrf2_trsm2_kernel, 256-thread CTA = two teams of 4 warps).
// ---------------------------------------------------------------------------------------------
constexpr int P2_THREADS = 256;

// C (64 x 64) -= A (64 x 64) B (64 x 64)^T, everything in shared memory (leading dimension LD2), by ONE team of 4 warps
// (tid in [0, 128)) with FP64 tensor MMAs; fragment layout as in gemm_tile_nt (warp tile 32 x 32).
__device__ __forceinline__ void smem_gemm64_nt_dmma(double* C, const double* A, const double* B, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
  const int g = lane >> 2, tg = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll 2
  for (int k0 = 0; k0 < NB; k0 += 4) {
    double fa[4], fb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) fa[a] = A[(wr + 8 * a + g) * LD2 + k0 + tg];
#pragma unroll
    for (int b = 0; b < 4; ++b) fb[b] = B[(wc + 8 * b + g) * LD2 + k0 + tg];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double2* p = reinterpret_cast<double2*>(C + (wr + 8 * a + g) * LD2 + wc + 8 * b + 2 * tg);
      double2 v = *p;
      v.x -= acc[a][b][0]; v.y -= acc[a][b][1];
      *p = v;
    }
}

// factor_solve_tile for a 256-thread CTA and NX (1 or 2) row tiles: the second row tile rides on warps 4-5 / 2-3 of the
// same phases (identical instruction streams side by side), its rank-32 update on the second team.
template <int NX>
__device__ __forceinline__ void factor_solve_tile2(double* sT, double* sX0, double* sX1, double* sLt, double* sinv, int* fail) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) potrf32_rl(sT, sLt, sinv, fail);
  __syncthreads();
  if (warp == 0) trsm32_row(sT + (HB + tid) * LD2, sLt, sinv);
  else if (warp == 2 || warp == 3) trsm32_row(sX0 + (tid - 64) * LD2, sLt, sinv);
  else if (NX == 2 && (warp == 4 || warp == 5)) trsm32_row(sX1 + (tid - 128) * LD2, sLt, sinv);
  __syncthreads();
  if (tid < 128) {
    gemm_nt32_il<2>(sT + HB * LD2 + HB, sT + HB * LD2, sT + HB * LD2, HB, tid);
    gemm_nt32_il<4>(sX0 + HB, sX0, sT + HB * LD2, NB, tid);
  } else if (NX == 2) {
    gemm_nt32_il<4>(sX1 + HB, sX1, sT + HB * LD2, NB, tid - 128);
  }
  __syncthreads();
  if (warp == 0) potrf32_rl(sT + HB * LD2 + HB, sLt + HB * LD2 + HB, sinv + HB, fail);
  __syncthreads();
  if (warp < 2) trsm32_row(sX0 + tid * LD2 + HB, sLt + HB * LD2 + HB, sinv + HB);
  else if (NX == 2 && warp < 4) trsm32_row(sX1 + (tid - 64) * LD2 + HB, sLt + HB * LD2 + HB, sinv + HB);
  __syncthreads();
}

}  // namespace tsl
