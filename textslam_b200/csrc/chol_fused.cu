// Fused reduced-system solve: the whole tile Cholesky of the reduced camera matrix, the forward solve riding on the b row and the
// backward solve as ONE persistent launch (replaces the ~21 wave launches of chol.cu: 10 x potrf_trsm_kernel, 9 x syrk_wave_kernel,
// backsolve_kernel and their grid-wide dependencies).
//
// Role in the reference: the linear solve inside ceres::Solve (src/optimizer.cc:1222,1602,1840), see chol.cu.
//
// CTAs pop tasks (chol_sched.hpp) from a topologically sorted queue and wait for their inputs on counters in global memory
// (release/acquire at gpu scope), so the trailing updates of one elimination level overlap the pivot chain of the next and the
// only serial part left is the chain of diagonal-block factorisations along the elimination tree:
//   F  one CTA factors a whole node (1 or 2 tiles = 2 or 4 blocks of 32 columns) in shared memory. The 32x32 diagonal blocks
//      are eliminated by ONE warp with every lane holding a full symmetric row in registers (potrf32_sym): the loop-carried chain
//      per pivot is reciprocal -> one FMA -> shuffle, and the same rank-1 updates that build the factor also build its inverse,
//      so everything off the diagonal is a GEMM against L^-1 on the FP64 tensor pipe (mma.sync.m8n8k4.f64; tcgen05 has no FP64
//      kind). The other seven warps run the trailing updates of the node and publish L_aa^-1 / L_ba while warp 0 is in its chain.
//   S  X = A_ij L_jj^-T for 8..64 rows of a tile (GEMM against the published inverse, half the k range by triangularity)
//   U  A_ik -= sum_j X_ij X_kj^T for one 32x32 quadrant, all source tiles of a wave in one pass
//   B  x_j = L_jj^-T (y_j - sum_i L_ij^T x_i): the L tiles are staged in shared memory before the wait for the x_i
// Data written by other CTAs of the launch is only ever read with ld.global.cg (L2), never through L1.
#include <algorithm>
#include "ctx.cuh"
#include "solver.cuh"
#include "chol_common.cuh"
#include "chol_sched.hpp"

namespace tsl {

constexpr int FTH = 256;             // threads per CTA
constexpr int LDB = 36;              // shared-memory stride of a 32x32 block (doubles): = 4 mod 16 -> conflict-free m8n8k4 fragment loads
constexpr int BS = HB * LDB;         // doubles per block
constexpr int LDW = NB + 4;          // stride of a 64-wide tile (= SPAD)
constexpr unsigned SPIN_LIMIT = 1u << 24;

struct FusedArgs {
  double* A; int ld; int Tn;
  const int* tasks; int ntasks;
  const int2* deps; const int* srcs; const int* below;
  int* sync;
  double* Linv; double* x; int* fail;
  unsigned long long* trace;   // optional: 4 stamps per task (pop, inputs ready, done, SM id)
};

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_s32(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 1/d to rounding level, branch-free: MUFU.RCP64H seed (PTX rcp.approx.ftz.f64, ~2^-20) + one cubic step y0 (1 + e + e^2)
__device__ __forceinline__ double rcp_pivot(double d) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-d, y0, 1.0);
  return fma(fma(e, e, e), y0, y0);
}

// acc (8x8 tile, m8n8k4 C fragment: c0 -> (row g, col 2 tg), c1 -> (row g, col 2 tg + 1)) += sum_{k0 <= k < k1} A[g][k] B[g][k];
// A points at the first row of an [m][k] operand, B at the first row of an [n][k] operand (both in shared memory).
__device__ __forceinline__ void mma_tile_nt(double& c0, double& c1, const double* A, int lda, const double* B, int ldb, int k0, int k1, int g, int tg) {
#pragma unroll 4
  for (int k = k0; k < k1; k += 4) dmma_m8n8k4(c0, c1, A[g * lda + k + tg], B[g * ldb + k + tg]);
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
__device__ __forceinline__ void potrf32_sym(const double* G, double* W, int* fail) {
  __shared__ __align__(16) double colbuf[2][HB];
  __shared__ double ssi[HB];
  const int r = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  __syncwarp();
  double v[HB];
#pragma unroll
  for (int k = 0; k < HB; ++k) v[k] = G[k * LDB + r];   // column r = row r
  // Software pipeline, one region per eliminated column between two warp barriers (ptxas schedules inside such a region only):
  // the region of column c applies column c with 1/d_c from the region before, and computes d_{c+1}, its reciprocal and the two
  // broadcast scalars of the next region while the 30-c update FMAs of every lane fill the stall slots of that chain.
  //   d  = pivot d_c (lane c's diagonal),  b = lane (c+1)'s entry of column c  (both by shuffle: they sit on the chain)
  //   colbuf[c & 1][k] = lane k's entry of column c (shared memory: feeds the FMAs, off the chain)
  bool bad = false;
  double d = __shfl_sync(full, v[0], 0);
  double b = __shfl_sync(full, v[0], 1);
  colbuf[0][r] = v[0];
  bad |= !(d > 0.0);
  d = d > 0.0 ? d : 1.0;
  double dr = d;                               // lane 0 keeps d_0; the others overwrite it at their own pivot
  double rinv = rcp_pivot(d);
  double vc = (r == 0) ? 0.0 : v[0];           // the pivot row itself is left alone
  double p1 = vc * b;
  __syncwarp();
#pragma unroll
  for (int c = 0; c + 1 < HB; ++c) {
    double col[HB];
#pragma unroll
    for (int k = (c + 2) & ~1; k < HB; k += 2) { const double2 t = *reinterpret_cast<const double2*>(&colbuf[c & 1][k]); col[k] = t.x; col[k + 1] = t.y; }
    v[c + 1] = fma(-p1, rinv, v[c + 1]);       // the only arithmetic between 1/d_c and d_{c+1}
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
  if (bad && r == 0) atomicExch(fail, 1);
  const double si = rsqrt_pivot(dr);
  ssi[r] = si;
  __syncwarp();
  const double nrr = -(si * si);   // -1 / d_r
#pragma unroll
  for (int k = 0; k < HB; ++k) {
    const double w = (k < r) ? 0.0 : ((k == r) ? si : ssi[k] * (nrr * v[k]));
    W[k * LDB + r] = w;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// F task
// ---------------------------------------------------------------------------------------------------------------------
// shared-memory layout of a node of nbk 32-column blocks: lower-triangular block storage G, one inverse per diagonal block,
// two column buffers of solved blocks X (even / odd columns alternate, so column h-1 is still there while column h is written)
__host__ __device__ constexpr int f_layout_doubles(int nbk) { return (nbk * (nbk + 1) / 2 + nbk + (nbk - 1) + (nbk >= 2 ? nbk - 2 : 0)) * BS; }
__device__ __forceinline__ double* f_blk(double* sG, int i, int j) { return sG + (i * (i + 1) / 2 + j) * BS; }

// L^-1 of a 64x64 tile from its two diagonal blocks' inverses Wp, Wq and the off-diagonal factor block Lqp:
//   [[Wp, 0], [-Wq (Lqp Wp), Wq]]   -> dst (64x64, row-major, tight). Executed by warps [w0, w0 + nw) ; tmp: one free block.
__device__ __noinline__ void linv_tile(const double* Wp, const double* Wq, const double* Lqp, double* tmp, double* dst, int w0, int nw, int bar_id) {
  const int warp = (threadIdx.x >> 5) - w0, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  // Mt[n][m] = (Lqp Wp)[m][n] = sum_{k >= n} Lqp[m][k] Wp[k][n]
  for (int u = warp; u < 16; u += nw) {
    const int ms = u >> 2, ns = u & 3;
    double c0 = 0.0, c1 = 0.0;
    for (int k = 8 * ns; k < HB; k += 4) dmma_m8n8k4(c0, c1, Lqp[(8 * ms + g) * LDB + k + tg], Wp[(k + tg) * LDB + 8 * ns + g]);
    tmp[(8 * ns + 2 * tg) * LDB + 8 * ms + g] = c0;
    tmp[(8 * ns + 2 * tg + 1) * LDB + 8 * ms + g] = c1;
  }
  // the three blocks that need no arithmetic
  const int t = threadIdx.x - 32 * w0, nt = 32 * nw;
  for (int e = t; e < HB * HB; e += nt) {
    const int rr = e >> 5, cc = e & 31;
    dst[rr * NB + cc] = Wp[rr * LDB + cc];
    dst[rr * NB + HB + cc] = 0.0;
    dst[(HB + rr) * NB + HB + cc] = Wq[rr * LDB + cc];
  }
  bar_named(bar_id, nt);
  for (int u = warp; u < 16; u += nw) {
    const int ms = u >> 2, ns = u & 3;
    double c0 = 0.0, c1 = 0.0;
    mma_tile_nt(c0, c1, Wq + 8 * ms * LDB, LDB, tmp + 8 * ns * LDB, LDB, 0, 8 * (ms + 1), g, tg);
    *reinterpret_cast<double2*>(dst + (HB + 8 * ms + g) * NB + 8 * ns + 2 * tg) = make_double2(-c0, -c1);
  }
}

// C (32x32 block) -= Xi Xk^T, tiles u = u0, u0 + ustep, ... of the 16 (8x8) tiles
__device__ __forceinline__ void blk_update(double* C, const double* Xi, const double* Xk, int u0, int ustep, int g, int tg) {
  for (int u = u0; u < 16; u += ustep) {
    const int ms = u >> 2, ns = u & 3;
    double c0 = 0.0, c1 = 0.0;
    mma_tile_nt(c0, c1, Xi + 8 * ms * LDB, LDB, Xk + 8 * ns * LDB, LDB, 0, HB, g, tg);
    double2* p = reinterpret_cast<double2*>(C + (8 * ms + g) * LDB + 8 * ns + 2 * tg);
    double2 v = *p; v.x -= c0; v.y -= c1; *p = v;
  }
}

// One node = nt tiles (1 or 2) = nbk = 2 nt blocks. Per block column h:  [T(h-1): X_i,h-1 = G_i,h-1 W_{h-1}^T for i >= h]
// [U1(h-1): G_hh -= X X^T]  then  warp 0: potrf32_sym(G_hh) -> W_h   |   warps 1..7: the rest of the rank-32 update of column
// h-1 (U2) and, once tile a is complete, the publication of L_ba and L_aa^-1 (so that the row tiles of column a and the updates
// they feed run on other SMs while this CTA factors tile b).
__device__ void f_task(const FusedArgs& a, const int* tk, double* smem) {
  const int nt = tk[FK_F_NT], nbk = 2 * nt, nblk = nbk * (nbk + 1) / 2;
  double* sG = smem;
  double* sW = sG + nblk * BS;
  double* sX0 = sW + nbk * BS;
  double* sX1 = sX0 + (nbk - 1) * BS;
  auto Xs = [&](int h, int i) { return ((h & 1) ? sX1 : sX0) + (i - h - 1) * BS; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int ta = tk[FK_F_TILE];
  const size_t ld = (size_t)a.ld;
  const double* An = a.A + (size_t)ta * NB * ld + (size_t)ta * NB;   // the node's diagonal super-tile
  // ---- load (diagonal blocks mirrored into both triangles) ----
#pragma unroll 4
  for (int e = tid; e < nblk * 512; e += FTH) {
    const int b = e >> 9, w = e & 511, rr = w >> 4, cc = (w & 15) * 2;
    const int bi = b >= 6 ? 3 : (b >= 3 ? 2 : (b >= 1 ? 1 : 0)), bj = b - bi * (bi + 1) / 2;
    const double2 val = ldcg2(An + (size_t)(HB * bi + rr) * ld + HB * bj + cc);
    double* dst = sG + b * BS;
    if (bi != bj) {
      *reinterpret_cast<double2*>(dst + rr * LDB + cc) = val;
    } else {
      if (cc <= rr) { dst[rr * LDB + cc] = val.x; dst[cc * LDB + rr] = val.x; }
      if (cc + 1 <= rr) { dst[rr * LDB + cc + 1] = val.y; dst[(cc + 1) * LDB + rr] = val.y; }
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int h = 0; h < nbk; ++h) {
    if (h > 0) {
      // ---- T(h-1): X_i = G_i,h-1 W_{h-1}^T for the blocks below the diagonal of column h-1 (k <= n by triangularity) ----
      const double* Wp = sW + (h - 1) * BS;
      const int nu = (nbk - h) * 16;
      for (int u = warp; u < nu; u += 8) {
        const int i = h + (u >> 4), ms = (u >> 2) & 3, ns = (u + (u >> 3)) & 3;
        double c0 = 0.0, c1 = 0.0;
        mma_tile_nt(c0, c1, f_blk(sG, i, h - 1) + 8 * ms * LDB, LDB, Wp + 8 * ns * LDB, LDB, 0, 8 * (ns + 1), g, tg);
        *reinterpret_cast<double2*>(Xs(h - 1, i) + (8 * ms + g) * LDB + 8 * ns + 2 * tg) = make_double2(c0, c1);
      }
      __syncthreads();
      // ---- U1(h-1): the next diagonal block (all 16 tiles: potrf32_sym wants both triangles) ----
      blk_update(f_blk(sG, h, h), Xs(h - 1, h), Xs(h - 1, h), warp, 8, g, tg);
      __syncthreads();
    }
    if (warp == 0) {
      potrf32_sym(f_blk(sG, h, h), sW + h * BS, a.fail);
    } else if (h > 0) {
      if (h == 2) {   // nt == 2: tile a is complete
        double* Lba = a.A + (size_t)(ta + 1) * NB * ld + (size_t)ta * NB;
        for (int e = tid - 32; e < 4 * 512; e += FTH - 32) {
          const int b = e >> 9, w = e & 511, rr = w >> 4, cc = (w & 15) * 2;
          const int hh = b & 1, i = 2 + (b >> 1);
          *reinterpret_cast<double2*>(Lba + (size_t)(HB * (i - 2) + rr) * ld + HB * hh + cc) = *reinterpret_cast<const double2*>(Xs(hh, i) + rr * LDB + cc);
        }
        linv_tile(sW, sW + BS, Xs(0, 1), f_blk(sG, 0, 0), a.Linv + (size_t)ta * NB * NB, 1, 7, 1);
        bar_named(1, FTH - 32);
        if (tid == 32) {
          __threadfence();
          red_release_add_s32(a.sync + tk[FK_F_XBA], 64);
          red_release_add_s32(a.sync + tk[FK_F_FIN], 1);
        }
      }
      // U2(h-1): remaining blocks (i, k), i >= k >= h, (i, k) != (h, h); tile u of a block belongs to warp 1 + (base + u) % 7
      int base = 0;
      for (int k = h; k < nbk; ++k)
        for (int i = (k == h ? h + 1 : k); i < nbk; ++i) {
          blk_update(f_blk(sG, i, k), Xs(h - 1, i), Xs(h - 1, k), (warp - 1 + 7 - (base % 7)) % 7, 7, g, tg);
          base += 16;
        }
    }
    __syncthreads();
  }
  // ---- last tile of the node: its inverse (X_qp = G_qp W_p^T is not formed by the loop: column nbk-2 has its T at h = nbk-1) ----
  {
    const int p = nbk - 2, q = nbk - 1;
    linv_tile(sW + p * BS, sW + q * BS, Xs(p, q), f_blk(sG, p, p), a.Linv + (size_t)(ta + nt - 1) * NB * NB, 0, 8, 2);
    __syncthreads();
    if (tid == 0) { __threadfence(); red_release_add_s32(a.sync + tk[FK_F_FIN] + nt - 1, 1); }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// S task: rows [row0, row0 + nrows) of X = A_ij L_jj^-T
// ---------------------------------------------------------------------------------------------------------------------
__device__ void s_task_run(const FusedArgs& a, const int* tk, double* smem) {
  double* sL = smem;                 // 64 x LDW: L_jj^-1 (row-major [n][k])
  double* sB = smem + NB * LDW;      // nrows x LDW
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int j = tk[FK_S_J], i = tk[FK_S_I], row0 = tk[FK_S_ROW0], nrows = tk[FK_S_NROWS];
  const size_t ld = (size_t)a.ld;
  const double* Lg = a.Linv + (size_t)j * NB * NB;
  double* Ag = a.A + ((size_t)i * NB + row0) * ld + (size_t)j * NB;
  for (int e = tid; e < NB * NB / 2; e += FTH) {
    const int rr = e >> 5, cc = (e & 31) * 2;
    *reinterpret_cast<double2*>(sL + rr * LDW + cc) = ldcg2(Lg + rr * NB + cc);
  }
  for (int e = tid; e < nrows * NB / 2; e += FTH) {
    const int rr = e >> 5, cc = (e & 31) * 2;
    *reinterpret_cast<double2*>(sB + rr * LDW + cc) = ldcg2(Ag + (size_t)rr * ld + cc);
  }
  __syncthreads();
  const int nu = (nrows >> 3) * 4;   // (8-row strip, pair of column tiles {c, 7 - c}): every unit has the same 18 k-steps
  for (int u = warp; u < nu; u += 8) {
    const int ms = u >> 2, cpair = u & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int ns = half ? 7 - cpair : cpair;
      double c0 = 0.0, c1 = 0.0;
      mma_tile_nt(c0, c1, sB + 8 * ms * LDW, LDW, sL + 8 * ns * LDW, LDW, 0, 8 * (ns + 1), g, tg);
      *reinterpret_cast<double2*>(Ag + (size_t)(8 * ms + g) * ld + 8 * ns + 2 * tg) = make_double2(c0, c1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// U task: 32x32 quadrant of A_ik -= sum_j X_ij X_kj^T
// ---------------------------------------------------------------------------------------------------------------------
__device__ void u_task(const FusedArgs& a, const int* tk, double* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int i = tk[FK_U_I], k = tk[FK_U_K], q = tk[FK_U_Q], qi = q >> 1, qk = q & 1;
  const int e0 = tk[FK_U_SRC0], e1 = tk[FK_U_SRC1];
  const size_t ld = (size_t)a.ld;
  const int ms = warp >> 1, ns0 = (warp & 1) * 2;
  const bool active = !(i == a.Tn && ms > 0);   // b row: only row 0 (first strip) carries data
  double* C = a.A + ((size_t)i * NB + HB * qi) * ld + (size_t)k * NB + HB * qk;
  double2 cv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) cv[t] = active ? ldcg2(C + (size_t)(8 * ms + g) * ld + 8 * (ns0 + t) + 2 * tg) : make_double2(0.0, 0.0);
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  double2 va[4], vb[4];
  auto fetch = [&](int j) {
    const double* Xi = a.A + ((size_t)i * NB + HB * qi) * ld + (size_t)j * NB;
    const double* Xk = a.A + ((size_t)k * NB + HB * qk) * ld + (size_t)j * NB;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + FTH * u, rr = e >> 5, cc = (e & 31) * 2;
      va[u] = ldcg2(Xi + (size_t)rr * ld + cc);
      vb[u] = ldcg2(Xk + (size_t)rr * ld + cc);
    }
  };
  if (e0 < e1) fetch(a.srcs[e0]);
  for (int e = e0; e < e1; ++e) {
    double* sA = smem + (e & 1) * 2 * HB * LDW;   // double buffered: one barrier per source
    double* sB = sA + HB * LDW;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ee = tid + FTH * u, rr = ee >> 5, cc = (ee & 31) * 2;
      *reinterpret_cast<double2*>(sA + rr * LDW + cc) = va[u];
      *reinterpret_cast<double2*>(sB + rr * LDW + cc) = vb[u];
    }
    __syncthreads();
    if (e + 1 < e1) fetch(a.srcs[e + 1]);
    if (active) {
#pragma unroll 4
      for (int k0 = 0; k0 < NB; k0 += 4) {
        const double fa = sA[(8 * ms + g) * LDW + k0 + tg];
#pragma unroll
        for (int t = 0; t < 2; ++t) dmma_m8n8k4(acc[t][0], acc[t][1], fa, sB[(8 * (ns0 + t) + g) * LDW + k0 + tg]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      double2 v = cv[t];
      v.x -= acc[t][0]; v.y -= acc[t][1];
      *reinterpret_cast<double2*>(C + (size_t)(8 * ms + g) * ld + 8 * (ns0 + t) + 2 * tg) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// waiting / signalling
// ---------------------------------------------------------------------------------------------------------------------
// Every thread takes some of the (sync index, minimum) pairs [d0, d1); returns false when the launch was aborted.
__device__ bool wait_deps(const FusedArgs& a, int d0, int d1, int* s_abort) {
  for (int e = d0 + (int)threadIdx.x; e < d1; e += FTH) {
    const int2 d = __ldg(a.deps + e);
    const int* p = a.sync + d.x;
    unsigned spins = 0;
    while (ld_acquire_s32(p) < d.y) {
      if ((++spins & 1023u) == 0) {
        if (ld_acquire_s32(a.sync + FS_ABORT) != 0) { *s_abort = 1; break; }
        if (spins > SPIN_LIMIT) { atomicExch(a.sync + FS_ABORT, 1); atomicExch(a.fail, 1); *s_abort = 1; break; }
      }
    }
  }
  __syncthreads();
  return *s_abort == 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// B task: x_j = L_jj^-T (y_j - sum_{i in below(j)} L_ij^T x_i)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int B_PRE = 4;   // L tiles staged in shared memory while the task waits for the x_i (the rest is read from L2 afterwards)
__device__ bool b_task(const FusedArgs& a, const int* tk, double* smem, int* s_abort) {
  __shared__ double sx[4][NB];
  __shared__ double st[4][NB];
  __shared__ double stt[NB];
  const int tid = threadIdx.x, c = tid & 63, g = tid >> 6;
  const int j = tk[FK_B_J], e0 = tk[FK_B_BEL0], e1 = tk[FK_B_BEL1];
  const size_t ld = (size_t)a.ld;
  double* sM = smem;                    // L_jj^-1, tight
  double* sT = smem + NB * NB;          // up to B_PRE tiles, tight
  // inputs that are final before any x_i is: L_jj^-1, the L_ij tiles, y_j (the first two dependencies + the xdone of the tiles)
  const int dmid = tk[FK_DEP0] + 1 + (e1 - e0) + 1;   // [fin_j, xdone(Tn, j), xdone(i, j)...] then [bx_i...]
  if (!wait_deps(a, tk[FK_DEP0], dmid, s_abort)) return false;
  {
    const double* M = a.Linv + (size_t)j * NB * NB;
    for (int e = tid; e < NB * NB / 2; e += FTH) *reinterpret_cast<double2*>(sM + 2 * e) = ldcg2(M + 2 * e);
    const int npre = min(e1 - e0, B_PRE);
    for (int t = 0; t < npre; ++t) {
      const double* L = a.A + (size_t)a.below[e0 + t] * NB * ld + (size_t)j * NB;
      for (int e = tid; e < NB * NB / 2; e += FTH) {
        const int rr = e >> 5, cc = (e & 31) * 2;
        *reinterpret_cast<double2*>(sT + t * NB * NB + rr * NB + cc) = ldcg2(L + (size_t)rr * ld + cc);
      }
    }
  }
  const double yj = (g == 0) ? __ldcg(a.A + (size_t)a.Tn * NB * ld + (size_t)j * NB + c) : 0.0;   // y_j = row 0 of the solved b tile
  if (!wait_deps(a, dmid, tk[FK_DEP1], s_abort)) return false;
  double t0 = 0.0, t1 = 0.0;
  for (int e = e0 + g; e < e1; e += 4) {
    const int i = a.below[e];
    sx[g][c] = __ldcg(a.x + (size_t)i * NB + c);
    bar_named(1 + g, 64);
    if (e - e0 < B_PRE) {
      const double* L = sT + (e - e0) * NB * NB + c;
#pragma unroll 16
      for (int r = 0; r < NB; r += 2) { t0 -= L[r * NB] * sx[g][r]; t1 -= L[(r + 1) * NB] * sx[g][r + 1]; }
    } else {
      const double* L = a.A + (size_t)i * NB * ld + (size_t)j * NB + c;
#pragma unroll 16
      for (int r = 0; r < NB; r += 2) { t0 -= __ldcg(L + (size_t)r * ld) * sx[g][r]; t1 -= __ldcg(L + (size_t)(r + 1) * ld) * sx[g][r + 1]; }
    }
    bar_named(1 + g, 64);
  }
  st[g][c] = t0 + t1;
  __syncthreads();
  if (g == 0) stt[c] = yj + ((st[0][c] + st[1][c]) + (st[2][c] + st[3][c]));
  __syncthreads();
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int r = 0; r < 16; r += 2) {
    a0 += sM[(16 * g + r) * NB + c] * stt[16 * g + r];
    a1 += sM[(16 * g + r + 1) * NB + c] * stt[16 * g + r + 1];
  }
  __syncthreads();
  st[g][c] = a0 + a1;
  __syncthreads();
  if (g == 0) a.x[(size_t)j * NB + c] = (st[0][c] + st[1][c]) + (st[2][c] + st[3][c]);
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FTH, 1) chol_fused_kernel(FusedArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task[F_TASK_INTS];
  __shared__ int s_next, s_abort;
  PDL_TRIGGER();
  const int tid = threadIdx.x;
  if (tid == 0) s_abort = 0;
  PDL_WAIT();
  for (;;) {
    __syncthreads();   // everybody is done with the previous task's shared memory and s_task
    if (tid == 0) s_next = atomicAdd(a.sync + FS_HEAD, 1);
    __syncthreads();
    const int t = s_next;
    if (t >= a.ntasks) break;
    if (tid < F_TASK_INTS) s_task[tid] = __ldg(a.tasks + (size_t)t * F_TASK_INTS + tid);
    __syncthreads();
    const int type = s_task[FK_TYPE];
    unsigned long long t_pop = 0, t_ready = 0;
    if (a.trace && tid == 0) t_pop = gtime();
    if (type == FT_B) {
      if (!b_task(a, s_task, smem, &s_abort)) break;
    } else {
      if (!wait_deps(a, s_task[FK_DEP0], s_task[FK_DEP1], &s_abort)) break;
      if (a.trace && tid == 0) t_ready = gtime();
      if (type == FT_S) s_task_run(a, s_task, smem);
      else if (type == FT_U) u_task(a, s_task, smem);
      else f_task(a, s_task, smem);
    }
    __syncthreads();   // all global stores of the task are issued
    if (tid == 0) {
      if (s_task[FK_SIG] >= 0) { __threadfence(); red_release_add_s32(a.sync + s_task[FK_SIG], s_task[FK_SIGINC]); }
      if (a.trace) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        a.trace[4 * (size_t)t] = t_pop; a.trace[4 * (size_t)t + 1] = t_ready; a.trace[4 * (size_t)t + 2] = gtime(); a.trace[4 * (size_t)t + 3] = smid;
      }
    }
  }
}

constexpr int FUSED_SMEM_DOUBLES = f_layout_doubles(4) > (1 + B_PRE) * NB * NB ? f_layout_doubles(4) : (1 + B_PRE) * NB * NB;
static_assert(FUSED_SMEM_DOUBLES * 8 <= 227 * 1024, "fused Cholesky exceeds the shared memory of one SM");
static_assert(2 * 2 * HB * LDW <= FUSED_SMEM_DOUBLES && 2 * NB * LDW <= FUSED_SMEM_DOUBLES, "S / U staging must fit");

__global__ void __launch_bounds__(256) zero_sync_kernel(int* __restrict__ sync, int n) {
  PDL_PROLOGUE();
  for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += gridDim.x * 256) sync[e] = 0;
}

int chol_fused_upload(tslam_ctx* ctx, const CholHost& H, CholSymbolic* sym) {
  cudaStream_t s = ctx->stream;
  sym->f_ntasks = H.f_ntasks; sym->f_nsync = H.f_nsync;
  static_assert(sizeof(I2) == sizeof(int2), "I2 must match int2");
  TSL_CUDA(sym->f_tasks.upload(H.f_tasks.data(), H.f_tasks.size(), s));
  TSL_CUDA(sym->f_deps.upload(reinterpret_cast<const int2*>(H.f_deps.data()), H.f_deps.size(), s));
  TSL_CUDA(sym->f_srcs.upload(H.f_srcs.data(), H.f_srcs.size(), s));
  TSL_CUDA(sym->f_below.upload(H.f_below.data(), H.f_below.size(), s));
  TSL_CUDA(sym->f_sync.reserve((size_t)(H.f_nsync ? H.f_nsync : 1)));
  return TSLAM_OK;
}

int chol_fused_clear(tslam_ctx* ctx, const CholSymbolic& sym) {
  LAUNCH(launch_k(zero_sync_kernel, (sym.f_nsync + 255) / 256, 256, 0, ctx->stream, sym.f_sync.p, sym.f_nsync));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

// Factor + both triangular solves; A as described in chol.cu (b row at tile row Tn), xout receives x. `sync` must have been
// cleared (chol_fused_clear) after the previous solve and before the reduced system was scattered.
int chol_fused_solve(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* xout, int* d_fail, unsigned long long* trace) {
  int ld, rows;
  const int Tn = chol_workspace_dims(sym.n, &ld, &rows);
  const int smem = FUSED_SMEM_DOUBLES * (int)sizeof(double);
  if (!ctx->attr_chol_fused) {
    TSL_CUDA(cudaFuncSetAttribute(chol_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ctx->attr_chol_fused = true;
  }
  FusedArgs a;
  a.A = A; a.ld = ld; a.Tn = Tn; a.tasks = sym.f_tasks.p; a.ntasks = sym.f_ntasks; a.deps = sym.f_deps.p; a.srcs = sym.f_srcs.p; a.below = sym.f_below.p;
  a.sync = sym.f_sync.p; a.Linv = sym.Ldiag.p; a.x = xout; a.fail = d_fail; a.trace = trace;
  const int grid = std::max(1, std::min(ctx->sm_count, sym.f_ntasks));
  LAUNCH(launch_k(chol_fused_kernel, grid, FTH, smem, ctx->stream, a));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl
