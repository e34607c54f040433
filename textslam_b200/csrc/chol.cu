// Dense-storage, tile-sparse FP64 Cholesky solve of the reduced camera system S y = b on one B200.
//
// Role in the reference: the linear solve inside ceres::Solve (src/optimizer.cc:1222,1602,1840) —
// the reference leaves Ceres' default (sparse normal Cholesky of the full system); eliminating the
// landmark blocks first and factoring the reduced camera matrix exactly is the same LM step
// (SURVEY Appendix A.7).
//
// Layout: A is row-major with leading dimension ld = Tn*64 (Tn = ceil(n/64)); rows [0,n) hold the
// lower triangle of S, rows [n, Tn*64) are identity padding, and one extra tile row starting at
// Rb = Tn*64 carries b in its first row. Factoring [S; b^T] panel by panel turns that row into
// y = L^-1 b for free (forward substitution rides along with the panel TRSM), so only the backward
// solve L^T x = y remains.
//
// Tile sparsity: the camera graph of a SLAM map is mostly banded (co-visibility), so most 64x64
// tiles of S and of its factor are structurally zero. The host does a symbolic factorisation on the
// Tn x Tn tile pattern once per problem (analysis.cpp: chol_symbolic_host) and every panel step only touches the listed
// non-zero tiles; a dense pattern degenerates to the classic right-looking blocked algorithm.
//
// Tile routines: fully unrolled, rows in registers (Crout). Measured alternatives (profiles/r1_notes.md): shared-memory
// left-looking loops (2.1x slower), shared-memory right-looking rank-1 updates on 256 threads (1.6x slower); the
// unrolled version is instruction-fetch bound (ncu: stall_no_instruction dominant, 70 % I-cache hit rate).
// Per 64-wide panel j: potrf (one CTA, rows in registers, Crout) -> trsm (one CTA per non-zero tile
// below) -> syrk/gemm trailing update with FP64 tensor-core MMA (mma.sync.m8n8k4.f64 — tcgen05 has
// no FP64 kind; DMMA is the FP64 tensor path on sm_100a), one CTA per non-zero 64x64 lower tile pair.
#include <algorithm>
#include "ctx.cuh"
#include "solver.cuh"

namespace tsl {

constexpr int NB = 64;
constexpr int SPAD = NB + 4;  // smem row stride (doubles): conflict-free m8n8k4 fragment loads

// ---------------------------------------------------------------------------------------------
// device tile routines for the 64x64 diagonal tile (CTA = 128 threads, tile and right-hand side in shared memory).
// The factorisation is recursive over 32x32 blocks: the sequential part (32-column Crout / substitution with the
// row held in registers, fully unrolled) exists ONCE as a __noinline__ function and is called twice, the coupling
// between the halves is a small register-tiled GEMM. A flat 64-column unrolled version measured 64 us per launch
// because 3 x 2016 FMAs of straight-line code miss the instruction cache (profiles/r1_notes.md).
// ---------------------------------------------------------------------------------------------
constexpr int HB = 32;             // half block
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

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
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

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// potrf of the diagonal tile fused with the TRSM of the non-zero tiles below it: every CTA factors
// the (tiny) diagonal tile redundantly into shared memory straight from A (nobody writes A_jj in
// this launch), CTA 0 stores the factor to Ldiag[j] (read by the backward solve), CTA 1+m solves the
// m-th non-zero row tile. Saves one launch + one grid-wide dependency per panel.
__global__ void __launch_bounds__(PT_THREADS) potrf_trsm_kernel(double* __restrict__ A, int ld, const int2* __restrict__ items,
                                                                int* __restrict__ fail, double* __restrict__ Linv) {
  // one CTA per (panel j, row tile i) item of the current wave; i < 0 marks the CTA that stores L_jj^-1.
  // Every CTA factors the (small) diagonal tile redundantly straight from A: nobody writes A_jj in this launch.
  extern __shared__ double smem[];
  double* sT = smem;                 // 64 x LDT: diagonal tile -> L_jj
  double* sX = smem + NB * LDT;      // 64 x LDT: row tile, or identity -> L_jj^-T
  __shared__ double sinv[NB];
  const int2 it = items[blockIdx.x];
  const int j = it.x;
  const double* Ajj = A + (size_t)j * NB * ld + (size_t)j * NB;
  double* Aij = it.y < 0 ? nullptr : A + (size_t)it.y * NB * ld + (size_t)j * NB;
#pragma unroll
  for (int half = 0; half < 2; ++half) {   // all global loads of a batch are issued before the first shared-memory store
    double vt[16], vx[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = threadIdx.x + PT_THREADS * (16 * half + u), r = e >> 6, c = e & 63;
      vt[u] = (c <= r) ? Ajj[(size_t)r * ld + c] : 0.0;
      vx[u] = Aij ? Aij[(size_t)r * ld + c] : ((r == c) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = threadIdx.x + PT_THREADS * (16 * half + u), r = e >> 6, c = e & 63;
      sT[r * LDT + c] = vt[u]; sX[r * LDT + c] = vx[u];
    }
  }
  __syncthreads();
  potrf_tile(sT, sinv, fail);
  __syncthreads();
  trsm_tile(sX, sT, sinv);             // row tile: X = A_ij L^-T ;  identity: X = L^-T
  if (!Aij) {                          // off the critical path: store L_jj^-1 = X^T (row-major) for the backward solve
    double* dst = Linv + (size_t)j * NB * NB;
    for (int e = threadIdx.x; e < NB * NB; e += PT_THREADS) { const int r = e >> 6, c = e & 63; dst[e] = sX[c * LDT + r]; }
    return;
  }
  for (int e = threadIdx.x; e < NB * NB; e += PT_THREADS) { const int r = e >> 6, c = e & 63; Aij[(size_t)r * ld + c] = sX[r * LDT + c]; }
}

// trailing update of one wave: one CTA per target tile (i,k); it sums the contributions X_i^(j) X_k^(j)^T of every
// source panel j of this wave (no two CTAs touch the same tile, so no atomics and a fixed summation order).
__global__ void __launch_bounds__(128) syrk_wave_kernel(double* __restrict__ A, int ld, const int2* __restrict__ targets,
                                                        const int* __restrict__ src_ptr, const int* __restrict__ src) {
  extern __shared__ double smem[];
  double* sA = smem;
  double* sB = smem + NB * SPAD;
  const int2 tg = targets[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
  const int g = lane >> 2, tgi = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int e = src_ptr[blockIdx.x]; e < src_ptr[blockIdx.x + 1]; ++e) {
    const int j = src[e];
    const double* Xi = A + (size_t)tg.x * NB * ld + (size_t)j * NB;
    const double* Xk = A + (size_t)tg.y * NB * ld + (size_t)j * NB;
    __syncthreads();   // previous source fully consumed
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double2 va[8], vb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int ee = threadIdx.x + 128 * (8 * half + u), r = ee >> 5, c2 = (ee & 31) * 2;
        va[u] = *reinterpret_cast<const double2*>(Xi + (size_t)r * ld + c2);
        vb[u] = *reinterpret_cast<const double2*>(Xk + (size_t)r * ld + c2);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int ee = threadIdx.x + 128 * (8 * half + u), r = ee >> 5, c2 = (ee & 31) * 2;
        sA[r * SPAD + c2] = va[u].x; sA[r * SPAD + c2 + 1] = va[u].y;
        sB[r * SPAD + c2] = vb[u].x; sB[r * SPAD + c2 + 1] = vb[u].y;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int k0 = 0; k0 < NB; k0 += 4) {
      double fa[4], fb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) fa[a] = sA[(wr + 8 * a + g) * SPAD + k0 + tgi];
#pragma unroll
      for (int b = 0; b < 4; ++b) fb[b] = sB[(wc + 8 * b + g) * SPAD + k0 + tgi];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
    }
  }
  double* C = A + (size_t)tg.x * NB * ld + (size_t)tg.y * NB;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = wr + 8 * a + g, c = wc + 8 * b + 2 * tgi;
      double2* p = reinterpret_cast<double2*>(C + (size_t)r * ld + c);
      double2 v = *p;
      v.x -= acc[a][b][0]; v.y -= acc[a][b][1];
      *p = v;
    }
}

// backward solve L^T x = y by waves in reverse order, gather form: for panel j of the wave
//   t = y_j - sum_{i > j, L_ij != 0} L_ij^T x_i   (x_i final: they belong to later waves)      x_j = L_jj^-T t
// x holds y on entry and is overwritten panel by panel. One CTA (64 threads = columns) per panel.
__global__ void __launch_bounds__(256) backsolve_wave_kernel(const double* __restrict__ A, int ld, const int* __restrict__ panels,
                                                             const int* __restrict__ below_ptr, const int* __restrict__ below,
                                                             const double* __restrict__ Linv, double* __restrict__ x) {
  // 256 threads = 4 groups x 64 columns; group g takes the tiles e = g (mod 4) of the list, partial sums meet in smem
  __shared__ double sx[4][NB];
  __shared__ double st[4][NB];
  __shared__ double stt[NB];
  const int j = panels[blockIdx.x];
  const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
  double t0 = 0.0, t1 = 0.0;
  const int e0 = below_ptr[blockIdx.x], e1 = below_ptr[blockIdx.x + 1];
  for (int e = e0 + g; e < e1; e += 4) {
    const int i = below[e];
    sx[g][c] = x[i * NB + c];
    __syncwarp();                                  // a group is 2 warps: make the tile's x visible with a named barrier
    asm volatile("bar.sync %0, 64;" ::"r"(g + 1));
    const double* Lij = A + (size_t)i * NB * ld + (size_t)j * NB + c;
#pragma unroll 8
    for (int r = 0; r < NB; r += 2) {
      t0 -= Lij[(size_t)r * ld] * sx[g][r];
      t1 -= Lij[(size_t)(r + 1) * ld] * sx[g][r + 1];
    }
    asm volatile("bar.sync %0, 64;" ::"r"(g + 1));  // everyone done with sx[g] before it is overwritten
  }
  st[g][c] = t0 + t1;
  __syncthreads();
  if (g == 0) stt[c] = x[j * NB + c] + ((st[0][c] + st[1][c]) + (st[2][c] + st[3][c]));
  __syncthreads();
  // x_c = sum_r (L^-1)[r][c] t_r, rows split over the 4 groups
  const double* M = Linv + (size_t)j * NB * NB;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
  for (int r = 16 * g; r < 16 * g + 16; r += 2) {
    a0 += M[r * NB + c] * stt[r];
    a1 += M[(r + 1) * NB + c] * stt[r + 1];
  }
  st[g][c] = a0 + a1;
  __syncthreads();
  if (g == 0) x[j * NB + c] = (st[0][c] + st[1][c]) + (st[2][c] + st[3][c]);
}

__global__ void copy_row_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// host: symbolic tile factorisation + launch sequence
// ---------------------------------------------------------------------------------------------
// Device copy of the tile-level symbolic factorisation + level schedule computed on the host (analysis.cpp:
// chol_symbolic_host): per wave the (panel, row tile) items of potrf_trsm_kernel, the target tiles of syrk_wave_kernel
// with their source panels, and the panel / below lists of the backward solve.
int chol_upload(tslam_ctx* ctx, const CholHost& H, CholSymbolic* sym) {
  sym->Tn = H.Tn; sym->n = H.n; sym->nwaves = H.nwaves; sym->gemm_tiles = H.gemm_tiles;
  sym->item_ptr = H.item_ptr; sym->target_ptr = H.target_ptr; sym->panel_ptr = H.panel_ptr;
  cudaStream_t s = ctx->stream;
  static_assert(sizeof(I2) == sizeof(int2), "I2 must match int2");
  TSL_CUDA(sym->items.upload(reinterpret_cast<const int2*>(H.items.data()), H.items.size(), s));
  TSL_CUDA(sym->targets.upload(reinterpret_cast<const int2*>(H.targets.data()), H.targets.size(), s));
  TSL_CUDA(sym->src_ptr.upload(H.src_ptr.data(), H.src_ptr.size(), s));
  TSL_CUDA(sym->src.upload(H.src.data(), H.src.size(), s));
  TSL_CUDA(sym->panels.upload(H.panels.data(), H.panels.size(), s));
  TSL_CUDA(sym->below_ptr.upload(H.below_ptr.data(), H.below_ptr.size(), s));
  TSL_CUDA(sym->below.upload(H.below.data(), H.below.size(), s));
  TSL_CUDA(sym->Ldiag.reserve((size_t)(H.Tn ? H.Tn : 1) * NB * NB));
  return TSLAM_OK;   // the caller synchronises the stream before H goes away
}

// Factor + solve. A: (Tn+1)*64 x ld as described above. xout: ld doubles (receives y, then x).
int chol_solve(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* ywork, double* xout, int* d_fail) {
  (void)ywork;
  int ld, rows;
  const int Tn = chol_workspace_dims(sym.n, &ld, &rows);
  static bool attr_set = false;
  const int smem = 2 * NB * SPAD * (int)sizeof(double);
  const int smem_pt = 2 * NB * LDT * (int)sizeof(double);
  if (!attr_set) {
    TSL_CUDA(cudaFuncSetAttribute(syrk_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TSL_CUDA(cudaFuncSetAttribute(potrf_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pt));
    attr_set = true;
  }
  cudaStream_t s = ctx->stream;
  for (int w = 0; w < sym.nwaves; ++w) {
    const int ni = sym.item_ptr[w + 1] - sym.item_ptr[w], nt = sym.target_ptr[w + 1] - sym.target_ptr[w];
    LAUNCH(potrf_trsm_kernel<<<ni, PT_THREADS, smem_pt, s>>>(A, ld, sym.items.p + sym.item_ptr[w], d_fail, sym.Ldiag.p));
    if (nt > 0) LAUNCH(syrk_wave_kernel<<<nt, 128, smem, s>>>(A, ld, sym.targets.p + sym.target_ptr[w], sym.src_ptr.p + sym.target_ptr[w], sym.src.p));
  }
  TSL_CHECK_LAUNCH();
  LAUNCH(copy_row_kernel<<<(ld + 255) / 256, 256, 0, s>>>(A + (size_t)Tn * NB * ld, xout, ld));
  for (int w = sym.nwaves - 1; w >= 0; --w) {
    const int np = sym.panel_ptr[w + 1] - sym.panel_ptr[w];
    LAUNCH(backsolve_wave_kernel<<<np, 256, 0, s>>>(A, ld, sym.panels.p + sym.panel_ptr[w], sym.below_ptr.p + sym.panel_ptr[w], sym.below.p,
                                                    sym.Ldiag.p, xout));
  }
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl
