// Dense-storage, tile-sparse FP64 Cholesky solve of the reduced camera system S y = b on one B200.
//
// Role in the reference: the linear solve inside ceres::Solve (src/optimizer.cc:1222,1602,1840) —
// the reference leaves Ceres' default (sparse normal Cholesky of the full system); eliminating the
// landmark blocks first and factoring the reduced camera matrix exactly is the same LM step
// (SURVEY Appendix A.7).
//
// Layout: A is row-major with leading dimension ld = Tn*64 (Tn = tiles of the tile-aligned camera layout, nd_layout.h); camera
// slot s owns rows / columns doff[s] .. doff[s]+5 of the lower triangle of S, every other row is identity padding (unit diagonal
// written by zero_tiles_kernel), and one extra tile row starting at Rb = Tn*64 carries b in its first row. Factoring [S; b^T] panel by panel turns that row into
// y = L^-1 b for free (forward substitution rides along with the panel TRSM), so only the backward
// solve L^T x = y remains.
//
// Tile sparsity: the camera graph of a SLAM map is mostly banded (co-visibility), so most 64x64
// tiles of S and of its factor are structurally zero. The host does a symbolic factorisation on the
// Tn x Tn tile pattern once per problem (analysis.cpp: chol_symbolic_host) and every panel step only touches the listed
// non-zero tiles; a dense pattern degenerates to the classic right-looking blocked algorithm.
//
// Per 64-wide panel j of a wave: potrf_trsm_kernel (one CTA per non-zero tile below the diagonal + one for L_jj^-1; every CTA
// re-factors the small diagonal tile, chol_tile.cuh: right-looking 32x32 factorisation inside one warp, rows in registers) ->
// syrk_wave_kernel, the trailing update with FP64 tensor-core MMA (mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind; DMMA is
// the FP64 tensor path on sm_100a), four CTAs per target tile. Waves come from the level schedule of the symbolic
// factorisation on the tile-aligned camera layout (nd_layout.h). potrf2_trsm2_kernel is the two-tile-panel variant
// (TSLAM_CHOL_PAIR=1). Measured alternatives and dead ends: profiles/r1_notes.md.
#include <algorithm>
#include "ctx.cuh"
#include "solver.cuh"
#include "chol_tile.cuh"
#include "chol_sched.hpp"

namespace tsl {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// potrf of the diagonal tile fused with the TRSM of the non-zero tiles below it: every CTA factors
// the (tiny) diagonal tile redundantly into shared memory straight from A (nobody writes A_jj in
// this launch), CTA 0 stores the factor to Ldiag[j] (read by the backward solve), CTA 1+m solves the
// m-th non-zero row tile. Saves one launch + one grid-wide dependency per panel.
template <bool MMA_UPDATE>
__global__ void __launch_bounds__(PT_THREADS) potrf_trsm_kernel(double* __restrict__ A, int ld, const int2* __restrict__ items,
                                                                int* __restrict__ fail, double* __restrict__ Linv) {
  PDL_TRIGGER();
  // one CTA per (panel j, row tile i) item of the current wave; i < 0 marks the CTA that stores L_jj^-1.
  // Every CTA factors the (small) diagonal tile redundantly straight from A: nobody writes A_jj in this launch.
  extern __shared__ __align__(16) double smem[];
  double* sT = smem;                   // 64 x LD2: diagonal tile -> L_jj
  double* sX = smem + NB * LD2;        // 64 x LD2: row tile, or identity -> L_jj^-T
  double* sLt = smem + 2 * NB * LD2;   // 64 x LD2: transposes of the two 32x32 diagonal blocks of L_jj
  __shared__ double sinv[NB];
  const int2 it = items[blockIdx.x];   // schedule of the symbolic factorisation: constant, read before the dependency wait
  PDL_WAIT();
  const int j = it.x;
  const double* Ajj = A + (size_t)j * NB * ld + (size_t)j * NB;
  double* Aij = it.y < 0 ? nullptr : A + (size_t)it.y * NB * ld + (size_t)j * NB;
#pragma unroll
  for (int half = 0; half < 2; ++half) {   // all global loads of a batch are issued before the first shared-memory store
    double2 vt[8], vx[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = threadIdx.x + PT_THREADS * (8 * half + u), r = e >> 5, c = (e & 31) * 2;
      vt[u] = *reinterpret_cast<const double2*>(Ajj + (size_t)r * ld + c);
      vx[u] = Aij ? *reinterpret_cast<const double2*>(Aij + (size_t)r * ld + c) : make_double2(r == c ? 1.0 : 0.0, r == c + 1 ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = threadIdx.x + PT_THREADS * (8 * half + u), r = e >> 5, c = (e & 31) * 2;
      *reinterpret_cast<double2*>(sT + r * LD2 + c) = make_double2(c <= r ? vt[u].x : 0.0, c + 1 <= r ? vt[u].y : 0.0);
      *reinterpret_cast<double2*>(sX + r * LD2 + c) = vx[u];
    }
  }
  __syncthreads();
  factor_solve_tile<false, MMA_UPDATE>(sT, sX, sLt, sinv, fail, nullptr);   // row tile: X = A_ij L^-T ;  identity: X = L^-T
  if (!Aij) {                          // off the critical path: store L_jj^-1 = X^T (row-major) for the backward solve
    double* dst = Linv + (size_t)j * NB * NB;
    for (int e = threadIdx.x; e < NB * NB; e += PT_THREADS) { const int r = e >> 6, c = e & 63; dst[e] = sX[c * LD2 + r]; }
    return;
  }
  for (int e = threadIdx.x; e < NB * NB / 2; e += PT_THREADS) {
    const int r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(Aij + (size_t)r * ld + c) = *reinterpret_cast<const double2*>(sX + r * LD2 + c);
  }
}

// Two-tile panel (a, b = a + 1): what two potrf_trsm launches and the update launch between them did for the tiles of one node
// of the camera layout, in ONE CTA per row tile (+ one for the diagonal):
//   L_aa = chol(A_aa);  L_ba = A_ba L_aa^-T;  X_a = B_a L_aa^-T          (factor_solve_tile2<2>)
//   A_bb -= L_ba L_ba^T;  B_b -= X_a L_ba^T                               (FP64 tensor MMAs, one team each)
//   L_bb = chol(A_bb);  X_b = B_b L_bb^-T                                  (factor_solve_tile2<1>)
// Every CTA redoes the (small) diagonal work straight from A: nobody writes A_aa, A_ba or A_bb in this launch. The diagonal
// CTA stores L_aa^-1, L_bb^-1 (backward solve) and parks L_ba in `Lpair` — copy_pair_tiles_kernel moves it into the factor
// once every CTA of the launch has read A_ba. Bit 30 of item.x: the row tile is structurally absent from column a.
__global__ void __launch_bounds__(P2_THREADS) potrf2_trsm2_kernel(double* __restrict__ A, int ld, const int2* __restrict__ items,
                                                                  int* __restrict__ fail, double* __restrict__ Linv, double* __restrict__ Lpair) {
  PDL_TRIGGER();
  extern __shared__ __align__(16) double smem[];
  double* sTa = smem;
  double* sXba = smem + 1 * NB * LD2;
  double* sXia = smem + 2 * NB * LD2;
  double* sTb = smem + 3 * NB * LD2;
  double* sXib = smem + 4 * NB * LD2;
  double* sLt = smem + 5 * NB * LD2;
  __shared__ double sinv[2 * NB];
  const int2 it = items[blockIdx.x];   // schedule of the symbolic factorisation: constant, read before the dependency wait
  PDL_WAIT();
  const int a = it.x & 0x3fffffff, b = a + 1, i = it.y;
  const bool has_a = (it.x & (1 << 30)) == 0, dg = i < 0;
  const double* Aaa = A + (size_t)a * NB * ld + (size_t)a * NB;
  const double* Aba = A + (size_t)b * NB * ld + (size_t)a * NB;
  const double* Abb = A + (size_t)b * NB * ld + (size_t)b * NB;
  double* Bia = dg ? nullptr : A + (size_t)i * NB * ld + (size_t)a * NB;
  double* Bib = dg ? nullptr : A + (size_t)i * NB * ld + (size_t)b * NB;
#pragma unroll
  for (int q = 0; q < 2; ++q) {   // all global loads of a batch are issued before the first shared-memory store
    double2 v0[4], v1[4], v2[4], v3[4], v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = threadIdx.x + P2_THREADS * (4 * q + u), r = e >> 5, c = (e & 31) * 2;
      const double2 idn = make_double2(r == c ? 1.0 : 0.0, r == c + 1 ? 1.0 : 0.0);
      v0[u] = *reinterpret_cast<const double2*>(Aaa + (size_t)r * ld + c);
      v1[u] = *reinterpret_cast<const double2*>(Aba + (size_t)r * ld + c);
      v2[u] = *reinterpret_cast<const double2*>(Abb + (size_t)r * ld + c);
      v3[u] = dg ? idn : (has_a ? *reinterpret_cast<const double2*>(Bia + (size_t)r * ld + c) : make_double2(0.0, 0.0));
      v4[u] = dg ? idn : *reinterpret_cast<const double2*>(Bib + (size_t)r * ld + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = threadIdx.x + P2_THREADS * (4 * q + u), r = e >> 5, c = (e & 31) * 2;
      *reinterpret_cast<double2*>(sTa + r * LD2 + c) = make_double2(c <= r ? v0[u].x : 0.0, c + 1 <= r ? v0[u].y : 0.0);
      *reinterpret_cast<double2*>(sXba + r * LD2 + c) = v1[u];
      *reinterpret_cast<double2*>(sTb + r * LD2 + c) = make_double2(c <= r ? v2[u].x : 0.0, c + 1 <= r ? v2[u].y : 0.0);
      *reinterpret_cast<double2*>(sXia + r * LD2 + c) = v3[u];
      *reinterpret_cast<double2*>(sXib + r * LD2 + c) = v4[u];
    }
  }
  __syncthreads();
  factor_solve_tile2<2>(sTa, sXba, sXia, sLt, sinv, fail);
  if (threadIdx.x < 128) smem_gemm64_nt_dmma(sTb, sXba, sXba, threadIdx.x);
  else if (!dg && has_a) smem_gemm64_nt_dmma(sXib, sXia, sXba, threadIdx.x - 128);
  __syncthreads();
  factor_solve_tile2<1>(sTb, sXib, nullptr, sLt, sinv + NB, fail);
  if (dg) {   // off the critical path: L_aa^-1 = X_a^T, L_bb^-1 = X_b^T (row-major) for the backward solve, L_ba for the factor
    double* da = Linv + (size_t)a * NB * NB;
    double* db = Linv + (size_t)b * NB * NB;
    double* dp = Lpair + (size_t)a * NB * NB;
    for (int e = threadIdx.x; e < NB * NB; e += P2_THREADS) {
      const int r = e >> 6, c = e & 63;
      da[e] = sXia[c * LD2 + r]; db[e] = sXib[c * LD2 + r]; dp[e] = sXba[r * LD2 + c];
    }
    return;
  }
  for (int e = threadIdx.x; e < NB * NB / 2; e += P2_THREADS) {
    const int r = e >> 5, c = (e & 31) * 2;
    if (has_a) *reinterpret_cast<double2*>(Bia + (size_t)r * ld + c) = *reinterpret_cast<const double2*>(sXia + r * LD2 + c);
    *reinterpret_cast<double2*>(Bib + (size_t)r * ld + c) = *reinterpret_cast<const double2*>(sXib + r * LD2 + c);
  }
}

// L_ba of every two-tile panel from its parking buffer into tile (a + 1, a) of the factor (read by the backward solve).
__global__ void __launch_bounds__(256) copy_pair_tiles_kernel(double* __restrict__ A, int ld, const int* __restrict__ pair_a, const double* __restrict__ Lpair) {
  PDL_TRIGGER();
  const int a = pair_a[blockIdx.x];
  PDL_WAIT();
  const double* src = Lpair + (size_t)a * NB * NB;
  double* dst = A + (size_t)(a + 1) * NB * ld + (size_t)a * NB;
  for (int e = threadIdx.x; e < NB * NB / 2; e += 256) {
    const int r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(dst + (size_t)r * ld + c) = *reinterpret_cast<const double2*>(src + r * NB + c);
  }
}

// trailing update of one wave: four CTAs per target tile (i,k), one per 32x32 quadrant; each sums the contributions
// X_i^(j) X_k^(j)^T of every source panel j of this wave (no two CTAs touch the same element, so no atomics and a fixed
// summation order). A wave has at most ~100 target tiles with 1-2 sources each, so a whole-tile CTA (64^3 FMAs = 2.2 us
// of one SM's FP64 pipe per source) left most SMs idle; quadrants quarter the critical path and fill the machine.
constexpr int QB = 32;
__global__ void __launch_bounds__(128) syrk_wave_kernel(double* __restrict__ A, int ld, const int2* __restrict__ targets,
                                                        const int* __restrict__ src_ptr, const int* __restrict__ src) {
  PDL_TRIGGER();
  extern __shared__ double smem[];
  double* sA = smem;               // 32 x SPAD: rows of X_i
  double* sB = smem + QB * SPAD;   // 32 x SPAD: rows of X_k
  const int t = blockIdx.x >> 2, qi = (blockIdx.x >> 1) & 1, qk = blockIdx.x & 1;
  const int2 tg = targets[t];
  const int e_begin = src_ptr[t], e_end = src_ptr[t + 1];
  int j_next = e_begin < e_end ? src[e_begin] : 0;   // the schedule is constant: read it before the dependency wait
  PDL_WAIT();
  if (tg.x == tg.y && qk > qi) return;   // the upper-right quadrant of a diagonal tile is never read
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp >> 1) * 16, wc = (warp & 1) * 16;
  const int g = lane >> 2, tgi = lane & 3;
  double acc[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  // the target quadrant is read now so that its latency hides behind the source loads and the MMA loop (no other CTA of
  // this launch writes it, and the previous launch has completed)
  double* C = A + ((size_t)tg.x * NB + QB * qi) * ld + (size_t)tg.y * NB + QB * qk;
  double2 cv[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) cv[a][b] = *reinterpret_cast<const double2*>(C + (size_t)(wr + 8 * a + g) * ld + wc + 8 * b + 2 * tgi);
  // software pipeline over the source panels: the 2 x 16 KB of source e + 1 are in flight (registers) while source e is
  // multiplied out of shared memory
  double2 va[8], vb[8];
  auto fetch = [&](int j) {
    const double* Xi = A + ((size_t)tg.x * NB + QB * qi) * ld + (size_t)j * NB;
    const double* Xk = A + ((size_t)tg.y * NB + QB * qk) * ld + (size_t)j * NB;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int ee = threadIdx.x + 128 * u, r = ee >> 5, c2 = (ee & 31) * 2;
      va[u] = *reinterpret_cast<const double2*>(Xi + (size_t)r * ld + c2);
      vb[u] = *reinterpret_cast<const double2*>(Xk + (size_t)r * ld + c2);
    }
  };
  if (e_begin < e_end) fetch(j_next);
  for (int e = e_begin; e < e_end; ++e) {
    __syncthreads();   // previous source fully consumed
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int ee = threadIdx.x + 128 * u, r = ee >> 5, c2 = (ee & 31) * 2;
      *reinterpret_cast<double2*>(sA + r * SPAD + c2) = va[u];
      *reinterpret_cast<double2*>(sB + r * SPAD + c2) = vb[u];
    }
    __syncthreads();
    if (e + 1 < e_end) fetch(src[e + 1]);
#pragma unroll 4
    for (int k0 = 0; k0 < NB; k0 += 4) {
      double fa[2], fb[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) fa[a] = sA[(wr + 8 * a + g) * SPAD + k0 + tgi];
#pragma unroll
      for (int b = 0; b < 2; ++b) fb[b] = sB[(wc + 8 * b + g) * SPAD + k0 + tgi];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int r = wr + 8 * a + g, c = wc + 8 * b + 2 * tgi;
      double2 v = cv[a][b];
      v.x -= acc[a][b][0]; v.y -= acc[a][b][1];
      *reinterpret_cast<double2*>(C + (size_t)r * ld + c) = v;
    }
}

// backward solve L^T x = y, gather form: for panel j
//   t = y_j - sum_{i > j, L_ij != 0} L_ij^T x_i          x_j = L_jj^-T t
// ONE launch for the whole solve: one CTA per panel, panels of the last wave first (blockIdx order), each CTA waits on
// the done-flags of the panels below it (they belong to later waves = lower block indices, so they are always scheduled
// no later than their consumers), and while it waits it already holds its first L_ij tile and its share of L_jj^-1 in
// registers. x holds y on entry and is overwritten panel by panel. flags[j] == epoch marks x_j final for this call.
__global__ void __launch_bounds__(256) backsolve_kernel(const double* __restrict__ A, int ld, int npanels, const int* __restrict__ panels,
                                                        const int* __restrict__ below_ptr, const int* __restrict__ below,
                                                        const double* __restrict__ Linv, const double* __restrict__ y, double* x, int* flags, int epoch) {
  PDL_TRIGGER();
  // 256 threads = 4 groups x 64 columns; group g takes the tiles e = g (mod 4) of the list, partial sums meet in smem
  __shared__ double sx[4][NB];
  __shared__ double st[4][NB];
  __shared__ double stt[NB];
  const int p = npanels - 1 - (int)blockIdx.x;
  const int j = panels[p];
  const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int e0 = below_ptr[p], e1 = below_ptr[p + 1];
  const int first_below = e0 + g < e1 ? below[e0 + g] : 0;   // schedule reads (constant) before the dependency wait
  PDL_WAIT();
  // ---- prefetch what does not depend on other panels ----
  double lfirst[NB];
  const bool has_first = e0 + g < e1;
  if (has_first) {
    const double* Lij = A + (size_t)first_below * NB * ld + (size_t)j * NB + c;
#pragma unroll
    for (int r = 0; r < NB; ++r) lfirst[r] = Lij[(size_t)r * ld];
  }
  double minv[16];
  {
    const double* M = Linv + (size_t)j * NB * NB;
#pragma unroll
    for (int r = 0; r < 16; ++r) minv[r] = M[(16 * g + r) * NB + c];
  }
  const double yj = (g == 0) ? y[j * NB + c] : 0.0;   // y_j = (L^-1 b)_j: the b row of the factorised workspace
  // ---- wait for the panels below ----
  for (int e = e0 + (int)threadIdx.x; e < e1; e += 256) {
    const volatile int* f = flags + below[e];
    while (*f != epoch) __nanosleep(32);
  }
  __threadfence();
  __syncthreads();
  double t0 = 0.0, t1 = 0.0;
  for (int e = e0 + g; e < e1; e += 4) {
    const int i = below[e];
    sx[g][c] = __ldcg(x + i * NB + c);             // written by another CTA of this launch: read through L2
    asm volatile("bar.sync %0, 64;" ::"r"(g + 1));  // a group is 2 warps: named barrier
    if (e == e0 + g) {
#pragma unroll
      for (int r = 0; r < NB; r += 2) { t0 -= lfirst[r] * sx[g][r]; t1 -= lfirst[r + 1] * sx[g][r + 1]; }
    } else {
      const double* Lij = A + (size_t)i * NB * ld + (size_t)j * NB + c;
#pragma unroll 16
      for (int r = 0; r < NB; r += 2) {
        t0 -= Lij[(size_t)r * ld] * sx[g][r];
        t1 -= Lij[(size_t)(r + 1) * ld] * sx[g][r + 1];
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(g + 1));  // everyone done with sx[g] before it is overwritten
  }
  st[g][c] = t0 + t1;
  __syncthreads();
  if (g == 0) stt[c] = yj + ((st[0][c] + st[1][c]) + (st[2][c] + st[3][c]));
  __syncthreads();
  // x_c = sum_r (L^-1)[r][c] t_r, rows split over the 4 groups
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int r = 0; r < 16; r += 2) {
    a0 += minv[r] * stt[16 * g + r];
    a1 += minv[r + 1] * stt[16 * g + r + 1];
  }
  st[g][c] = a0 + a1;
  __syncthreads();
  if (g == 0) {
    x[j * NB + c] = (st[0][c] + st[1][c]) + (st[2][c] + st[3][c]);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); *reinterpret_cast<volatile int*>(flags + j) = epoch; }
}

// Zero the tiles of the factor pattern (one CTA per (panel, row tile) item = one 64x64 tile) before the reduced system is
// scattered into them: the dense workspace is 10x larger than its structurally non-zero part on a SLAM camera graph.
__global__ void __launch_bounds__(256) zero_tiles_kernel(double* __restrict__ A, int ld, const int2* __restrict__ items) {
  PDL_PROLOGUE();
  const int2 it = items[blockIdx.x];
  double* T = A + (size_t)(it.y < 0 ? it.x : it.y) * NB * ld + (size_t)it.x * NB;
  // diagonal tiles get a unit diagonal: the columns no camera owns (padding of the tile-aligned layout, nd_layout.h) stay
  // decoupled identity rows, the real diagonal entries are overwritten by scatter_kernel
  const bool dg = it.y < 0;
  for (int e = threadIdx.x; e < NB * NB / 2; e += 256) {
    const int r = e >> 5, c = (e & 31) * 2;
    *reinterpret_cast<double2*>(T + (size_t)r * ld + c) = make_double2(dg && r == c ? 1.0 : 0.0, dg && r == c + 1 ? 1.0 : 0.0);
  }
}

__global__ void copy_row_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

bool chol_fused_enabled() {
  static const bool on = [] { const char* e = getenv("TSLAM_CHOL_FUSED"); return !(e && e[0] == '0'); }();
  return on;
}

// ---------------------------------------------------------------------------------------------
// host: symbolic tile factorisation + launch sequence
// ---------------------------------------------------------------------------------------------
// Device copy of the tile-level symbolic factorisation + level schedule computed on the host (analysis.cpp:
// chol_symbolic_host): per wave the (panel, row tile) items of potrf_trsm_kernel, the target tiles of syrk_wave_kernel
// with their source panels, and the panel / below lists of the backward solve.
int chol_upload(tslam_ctx* ctx, const CholHost& H, CholSymbolic* sym) {
  sym->Tn = H.Tn; sym->n = H.n; sym->nwaves = H.nwaves; sym->gemm_tiles = H.gemm_tiles;
  sym->item_ptr.assign(H.item_ptr.begin(), H.item_ptr.end()); sym->target_ptr.assign(H.target_ptr.begin(), H.target_ptr.end());
  sym->item2_ptr.assign(H.item2_ptr.begin(), H.item2_ptr.end()); sym->n_clear = (int)H.clear_items.size();
  sym->panel_ptr.assign(H.panel_ptr.begin(), H.panel_ptr.end());
  cudaStream_t s = ctx->stream;
  static_assert(sizeof(I2) == sizeof(int2), "I2 must match int2");
  TSL_CUDA(sym->items.upload(reinterpret_cast<const int2*>(H.items.data()), H.items.size(), s));
  TSL_CUDA(sym->items2.upload(reinterpret_cast<const int2*>(H.items2.data()), H.items2.size(), s));
  TSL_CUDA(sym->clear_items.upload(reinterpret_cast<const int2*>(H.clear_items.data()), H.clear_items.size(), s));
  TSL_CUDA(sym->targets.upload(reinterpret_cast<const int2*>(H.targets.data()), H.targets.size(), s));
  TSL_CUDA(sym->src_ptr.upload(H.src_ptr.data(), H.src_ptr.size(), s));
  TSL_CUDA(sym->src.upload(H.src.data(), H.src.size(), s));
  TSL_CUDA(sym->panels.upload(H.panels.data(), H.panels.size(), s));
  TSL_CUDA(sym->below_ptr.upload(H.below_ptr.data(), H.below_ptr.size(), s));
  TSL_CUDA(sym->below.upload(H.below.data(), H.below.size(), s));
  TSL_CUDA(sym->Ldiag.reserve((size_t)(H.Tn ? H.Tn : 1) * NB * NB));
  {
    std::vector<int> pa;
    for (const I2& it : H.items2) if (it.y < 0) pa.push_back(it.x);
    sym->n_pairs = (int)pa.size();
    TSL_CUDA(sym->pair_a.upload(pa.data(), pa.size(), s));
    TSL_CUDA(cudaStreamSynchronize(s));   // pa is a local
    if (sym->n_pairs) TSL_CUDA(sym->Lpair.reserve((size_t)H.Tn * NB * NB));
  }
  TSL_CUDA(sym->flags.reserve((size_t)(H.Tn ? H.Tn : 1)));
  TSL_CUDA(cudaMemsetAsync(sym->flags.p, 0, sizeof(int) * (size_t)(H.Tn ? H.Tn : 1), s));
  sym->epoch = 0;
  { int rc = chol_fused_upload(ctx, H, sym); if (rc) return rc; }
  return TSLAM_OK;   // the caller synchronises the stream before H goes away
}

// Clears every tile the factorisation reads or writes (pattern of L incl. the b row); everything else is never touched.
int chol_clear(tslam_ctx* ctx, const CholSymbolic& sym, double* A) {
  int ld, rows;
  chol_workspace_dims(sym.n, &ld, &rows);
  const int ni = sym.n_clear;
  if (ni > 0) LAUNCH(launch_k(zero_tiles_kernel, ni, 256, 0, ctx->stream, A, ld, sym.clear_items.p));
  TSL_CHECK_LAUNCH();
  if (chol_fused_enabled() && sym.f_ntasks > 0) return chol_fused_clear(ctx, sym);
  return TSLAM_OK;
}

// Factor + solve. A: (Tn+1)*64 x ld as described above. xout: ld doubles (receives y, then x).
int chol_solve(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* ywork, double* xout, int* d_fail) {
  (void)ywork;
  if (chol_fused_enabled() && sym.f_ntasks > 0) return chol_fused_solve(ctx, sym, A, xout, d_fail, nullptr);
  return chol_solve_waves(ctx, sym, A, xout, d_fail);
}

// The wave-scheduled launch sequence (TSLAM_CHOL_FUSED=0): one potrf_trsm + one syrk launch per wave, then the backward solve.
int chol_solve_waves(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* xout, int* d_fail) {
  int ld, rows;
  const int Tn = chol_workspace_dims(sym.n, &ld, &rows);
  const int smem = 2 * QB * SPAD * (int)sizeof(double);
  const int smem_pt = 3 * NB * LD2 * (int)sizeof(double);
  const int smem_p2 = 6 * NB * LD2 * (int)sizeof(double);
  if (!ctx->attr_chol_waves) {
    TSL_CUDA(cudaFuncSetAttribute(potrf2_trsm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p2));
    TSL_CUDA(cudaFuncSetAttribute(syrk_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TSL_CUDA(cudaFuncSetAttribute(potrf_trsm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pt));
    TSL_CUDA(cudaFuncSetAttribute(potrf_trsm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pt));
    ctx->attr_chol_waves = true;
  }
  cudaStream_t s = ctx->stream;
  // TSLAM_CHOL_TRACE=1: per-kernel-class device time of this call (CUDA events between launches; debugging aid only)
  static const bool trace = getenv("TSLAM_CHOL_TRACE") != nullptr;
  // in-tile rank-32 updates of the panel kernel on the FP64 tensor pipe (TSLAM_POTRF_FMA=1: the register-tiled FMA version)
  static const bool mma_update = getenv("TSLAM_POTRF_FMA") == nullptr;
  std::vector<cudaEvent_t> ev; std::vector<int> cls;
  auto mark = [&](int c) { if (!trace) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); ev.push_back(e); cls.push_back(c); };
  mark(-1);
  for (int w = 0; w < sym.nwaves; ++w) {
    const int ni = sym.item_ptr[w + 1] - sym.item_ptr[w], ni2 = sym.item2_ptr[w + 1] - sym.item2_ptr[w], nt = sym.target_ptr[w + 1] - sym.target_ptr[w];
    if (ni2 > 0) LAUNCH(launch_k(potrf2_trsm2_kernel, ni2, P2_THREADS, smem_p2, s, A, ld, sym.items2.p + sym.item2_ptr[w], d_fail, sym.Ldiag.p, sym.Lpair.p));
    if (ni > 0) {
      if (mma_update) LAUNCH(launch_k(potrf_trsm_kernel<true>, ni, PT_THREADS, smem_pt, s, A, ld, sym.items.p + sym.item_ptr[w], d_fail, sym.Ldiag.p));
      else LAUNCH(launch_k(potrf_trsm_kernel<false>, ni, PT_THREADS, smem_pt, s, A, ld, sym.items.p + sym.item_ptr[w], d_fail, sym.Ldiag.p));
    }
    mark(0);
    if (nt > 0) LAUNCH(launch_k(syrk_wave_kernel, 4 * nt, 128, smem, s, A, ld, sym.targets.p + sym.target_ptr[w], sym.src_ptr.p + sym.target_ptr[w], sym.src.p));
    mark(1);
  }
  TSL_CHECK_LAUNCH();
  if (sym.n_pairs > 0) LAUNCH(launch_k(copy_pair_tiles_kernel, sym.n_pairs, 256, 0, s, A, ld, sym.pair_a.p, sym.Lpair.p));
  {
    const int np = sym.panel_ptr[sym.nwaves];
    const int epoch = ++sym.epoch;
    LAUNCH(launch_k(backsolve_kernel, np, 256, 0, s, A, ld, np, sym.panels.p, sym.below_ptr.p, sym.below.p, sym.Ldiag.p, A + (size_t)Tn * NB * ld, xout, sym.flags.p, epoch));
    mark(3);
  }
  if (trace) {
    cudaStreamSynchronize(s);
    float tot[4] = {0, 0, 0, 0};
    for (size_t k = 1; k < ev.size(); ++k) { float ms = 0; cudaEventElapsedTime(&ms, ev[k - 1], ev[k]); tot[cls[k]] += ms; }
    fprintf(stderr, "[tslam chol] waves %d: potrf_trsm %.1f us, syrk %.1f us, copy %.1f us, backsolve %.1f us\n", sym.nwaves, tot[0] * 1e3, tot[1] * 1e3, tot[2] * 1e3, tot[3] * 1e3);
    for (auto e : ev) cudaEventDestroy(e);
  }
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl

using namespace tsl;

// Test hook (host only, no device needed): the fused task schedule of an n x n system with the given lower-triangular tile
// pattern (Tn x Tn flags, Tn = ceil(n / 64)). counts_out = {tasks, deps, srcs, below, sync ints, Tn}; the arrays are filled up to
// their capacities (call once with zero capacities for the sizes).
extern "C" int tslam_debug_chol_schedule(int n, const uint8_t* tile_nz, int32_t counts_out[6], int32_t* tasks, int cap_tasks, int32_t* deps, int cap_deps,
                                         int32_t* srcs, int cap_srcs, int32_t* below, int cap_below) {
  if (n <= 0 || !tile_nz || !counts_out) return set_error(TSLAM_ERR_ARG, "bad argument");
  Arena arena([](size_t b) -> void* { return malloc(b); }, [](void* q) { free(q); });
  CholHost H;
  try { chol_symbolic_in_arena(n, tile_nz, arena, H); } catch (const std::exception& e) { return set_error(TSLAM_ERR_ARG, "symbolic factorisation failed: %s", e.what()); }
  counts_out[0] = H.f_ntasks; counts_out[1] = (int)H.f_deps.size(); counts_out[2] = (int)H.f_srcs.size(); counts_out[3] = (int)H.f_below.size();
  counts_out[4] = H.f_nsync; counts_out[5] = H.Tn;
  if (tasks) std::copy(H.f_tasks.begin(), H.f_tasks.begin() + std::min<size_t>(H.f_tasks.size(), (size_t)cap_tasks * F_TASK_INTS), tasks);
  if (deps) for (size_t e = 0; e < H.f_deps.size() && e < (size_t)cap_deps; ++e) { deps[2 * e] = H.f_deps[e].x; deps[2 * e + 1] = H.f_deps[e].y; }
  if (srcs) std::copy(H.f_srcs.begin(), H.f_srcs.begin() + std::min<size_t>(H.f_srcs.size(), (size_t)cap_srcs), srcs);
  if (below) std::copy(H.f_below.begin(), H.f_below.begin() + std::min<size_t>(H.f_below.size(), (size_t)cap_below), below);
  return TSLAM_OK;
}

// Test / bench hook: solves S x = b for a dense symmetric positive definite S (n x n, row-major, lower triangle read) through the
// reduced-system solver alone. tile_nz: Tn x Tn lower tile pattern or NULL (derived from the non-zeros of S).
// mode 0 = wave kernels, 1 = fused persistent kernel. ms_out = mean device time of the solve over `reps` runs (the workspace is
// restored before each). trace_out (fused only): 16 uint64 per task (pop, inputs ready, done [ns], SM id, then clock64 phase stamps of F tasks), up to trace_cap tasks.
extern "C" int tslam_dev_chol_solve(tslam_ctx* ctx, int n, const uint8_t* tile_nz, const double* S, const double* b, double* x_out, int mode, int reps,
                                    float* ms_out, uint64_t* trace_out, int trace_cap, int32_t* info_out /*[4]: tasks, waves, Tn, fail*/) {
  if (!ctx || n <= 0 || !S || !b || !x_out) return set_error(TSLAM_ERR_ARG, "null argument");
  TSL_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int ld, rows;
  const int Tn = chol_workspace_dims(n, &ld, &rows);
  std::vector<uint8_t> tz((size_t)Tn * Tn, 0);
  if (tile_nz) tz.assign(tile_nz, tile_nz + (size_t)Tn * Tn);
  else
    for (int r = 0; r < n; ++r) for (int c = 0; c <= r; ++c) if (S[(size_t)r * n + c] != 0.0) tz[(size_t)(r / NB) * Tn + c / NB] = 1;
  if (!ctx->host_arena)
    ctx->host_arena = new Arena([](size_t nb) -> void* { void* q = nullptr; return cudaHostAlloc(&q, nb, cudaHostAllocDefault) == cudaSuccess ? q : nullptr; },
                                [](void* q) { cudaFreeHost(q); });
  CholHost H;
  try { chol_symbolic_in_arena(n, tz.data(), *ctx->host_arena, H); } catch (const std::exception& e) { return set_error(TSLAM_ERR_CUDA, "symbolic factorisation failed: %s", e.what()); }
  CholSymbolic sym;
  int rc = chol_upload(ctx, H, &sym);
  if (rc) return rc;
  // padded workspace image: lower triangle of S, unit diagonal on the padding, b in the first row of the extra tile row
  std::vector<double> W0((size_t)rows * ld, 0.0);
  for (int r = 0; r < n; ++r) for (int c = 0; c <= r; ++c) W0[(size_t)r * ld + c] = S[(size_t)r * n + c];
  for (int r = n; r < Tn * NB; ++r) W0[(size_t)r * ld + r] = 1.0;
  for (int c = 0; c < n; ++c) W0[(size_t)Tn * NB * ld + c] = b[c];
  DevBuf<double> A0, A, x; DevBuf<int> fail; DevBuf<unsigned long long> trace;
  TSL_CUDA(A0.upload(W0.data(), W0.size(), st)); TSL_CUDA(A.reserve(W0.size())); TSL_CUDA(x.reserve(ld)); TSL_CUDA(fail.reserve(1));
  TSL_CUDA(cudaMemsetAsync(fail.p, 0, sizeof(int), st));
  const bool fused = mode == 1;
  if (fused && trace_out) { TSL_CUDA(trace.reserve(16 * (size_t)H.f_ntasks)); TSL_CUDA(cudaMemsetAsync(trace.p, 0, 128 * (size_t)H.f_ntasks, st)); }
  float total = 0.f;
  for (int it = 0; it < std::max(1, reps); ++it) {
    TSL_CUDA(cudaMemcpyAsync(A.p, A0.p, W0.size() * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (fused) { if ((rc = chol_fused_clear(ctx, sym))) return rc; }
    TSL_CUDA(cudaEventRecord(ctx->ev0, st));
    if (fused) rc = chol_fused_solve(ctx, sym, A.p, x.p, fail.p, trace_out ? trace.p : nullptr);
    else rc = chol_solve_waves(ctx, sym, A.p, x.p, fail.p);
    if (rc) return rc;
    TSL_CUDA(cudaEventRecord(ctx->ev1, st));
    TSL_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    TSL_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    total += ms;
  }
  if (ms_out) *ms_out = total / std::max(1, reps);
  std::vector<double> xh(ld);
  int hfail = 0;
  TSL_CUDA(cudaMemcpyAsync(xh.data(), x.p, sizeof(double) * ld, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(&hfail, fail.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (fused && trace_out) TSL_CUDA(cudaMemcpyAsync(trace_out, trace.p, 128 * (size_t)std::min(trace_cap, H.f_ntasks), cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  std::copy(xh.begin(), xh.begin() + n, x_out);
  if (info_out) { info_out[0] = H.f_ntasks; info_out[1] = H.nwaves; info_out[2] = Tn; info_out[3] = hfail; }
  return TSLAM_OK;
}
