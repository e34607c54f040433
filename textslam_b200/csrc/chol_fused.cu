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
#include "chol_potrf.cuh"
#include "chol_sched.hpp"

namespace tsl {

constexpr int FTH = 256;             // threads per CTA
constexpr int BS = HB * LDB;         // doubles per block
constexpr int LDW = NB + 4;          // stride of a 64-wide tile (= SPAD)
constexpr unsigned SPIN_LIMIT = 1u << 24;
constexpr int TRACE_WORDS = 16;   // uint64 per task in the optional trace

struct FusedArgs {
  double* A; int ld; int Tn;
  const int* tasks; int ntasks;
  const int2* deps; const int2* dep_inl; const int* idx_inl; const int* srcs; const int* below;   // dep_inl: first F_INL pairs of every task, queue order
  int* sync;
  double* Linv; double* x; int* fail;
  unsigned long long* trace;   // optional: 4 stamps per task (pop, inputs ready, done, SM id) + 12 phase stamps (clock64) of F tasks
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
// 16-byte asynchronous copy global -> shared through L2 only (.cg: coherent with what other SMs of this launch have released);
// no register staging, so a task can have its whole input in flight at once
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// acc (8x8 tile, m8n8k4 C fragment: c0 -> (row g, col 2 tg), c1 -> (row g, col 2 tg + 1)) += sum_{k0 <= k < k1} A[g][k] B[g][k];
// A points at the first row of an [m][k] operand, B at the first row of an [n][k] operand (both in shared memory).
__device__ __forceinline__ void mma_tile_nt(double& c0, double& c1, const double* A, int lda, const double* B, int ldb, int k0, int k1, int g, int tg) {
#pragma unroll 4
  for (int k = k0; k < k1; k += 4) dmma_m8n8k4(c0, c1, A[g * lda + k + tg], B[g * ldb + k + tg]);
}

// ---------------------------------------------------------------------------------------------------------------------
// F task
// ---------------------------------------------------------------------------------------------------------------------
// shared-memory layout of a node of nbk 32-column blocks: lower-triangular block storage G, one inverse per diagonal block,
// two column buffers of solved blocks X (even / odd columns alternate, so column h-1 is still there while column h is written)
__host__ __device__ constexpr int f_layout_doubles(int nbk) { return (nbk * (nbk + 1) / 2 + nbk + (nbk - 1) + (nbk >= 2 ? nbk - 2 : 0)) * BS; }
__device__ __forceinline__ double* f_blk(double* sG, int i, int j) { return sG + (i * (i + 1) / 2 + j) * BS; }

// two 8x8 tiles that share the A rows: c[0..1] += A B0^T, c[2..3] += A B1^T over k0 <= k < k1 (two independent MMA chains per warp)
__device__ __forceinline__ void mma_pair_nt(double c[4], const double* A, int lda, const double* B0, const double* B1, int ldb, int k0, int k1, int g, int tg) {
#pragma unroll 4
  for (int k = k0; k < k1; k += 4) {
    const double fa = A[g * lda + k + tg];
    dmma_m8n8k4(c[0], c[1], fa, B0[g * ldb + k + tg]);
    dmma_m8n8k4(c[2], c[3], fa, B1[g * ldb + k + tg]);
  }
}

// L^-1 of a 64x64 tile from its two diagonal blocks' inverses Wp, Wq and the off-diagonal factor block Lqp:
//   [[Wp, 0], [-Wq (Lqp Wp), Wq]]   -> dst (64x64, row-major, tight). Executed by nw warps (this warp is number wi of them,
// this thread number ti of 32 nw); tmp: one free block; bar_id: a named barrier for exactly these warps.
__device__ __noinline__ void linv_tile(const double* Wp, const double* Wq, const double* Lqp, double* tmp, double* dst, int wi, int nw, int ti, int bar_id) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  // Mt[n][m] = (Lqp Wp)[m][n] = sum_{k >= n} Lqp[m][k] Wp[k][n]; unit = two row strips of one column tile (two MMA chains)
  for (int u = wi; u < 8; u += nw) {
    const int ns = u & 3, ms0 = (u >> 2) * 2;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 8 * ns; k < HB; k += 4) {
      const double fb = Wp[(k + tg) * LDB + 8 * ns + g];
      dmma_m8n8k4(c[0], c[1], Lqp[(8 * ms0 + g) * LDB + k + tg], fb);
      dmma_m8n8k4(c[2], c[3], Lqp[(8 * ms0 + 8 + g) * LDB + k + tg], fb);
    }
    tmp[(8 * ns + 2 * tg) * LDB + 8 * ms0 + g] = c[0];
    tmp[(8 * ns + 2 * tg + 1) * LDB + 8 * ms0 + g] = c[1];
    tmp[(8 * ns + 2 * tg) * LDB + 8 * ms0 + 8 + g] = c[2];
    tmp[(8 * ns + 2 * tg + 1) * LDB + 8 * ms0 + 8 + g] = c[3];
  }
  // the three blocks that need no arithmetic
  const int nt = 32 * nw;
  for (int e = ti; e < HB * HB; e += nt) {
    const int rr = e >> 5, cc = e & 31;
    dst[rr * NB + cc] = Wp[rr * LDB + cc];
    dst[rr * NB + HB + cc] = 0.0;
    dst[(HB + rr) * NB + HB + cc] = Wq[rr * LDB + cc];
  }
  bar_named(bar_id, nt);
  for (int u = wi; u < 8; u += nw) {   // unit = two column tiles of one row strip
    const int ms = u >> 1, ns0 = (u & 1) * 2;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    mma_pair_nt(c, Wq + 8 * ms * LDB, LDB, tmp + 8 * ns0 * LDB, tmp + 8 * (ns0 + 1) * LDB, LDB, 0, 8 * (ms + 1), g, tg);
    *reinterpret_cast<double2*>(dst + (HB + 8 * ms + g) * NB + 8 * ns0 + 2 * tg) = make_double2(-c[0], -c[1]);
    *reinterpret_cast<double2*>(dst + (HB + 8 * ms + g) * NB + 8 * ns0 + 8 + 2 * tg) = make_double2(-c[2], -c[3]);
  }
}

// C (32x32 block) -= Xi Xk^T: the 8 tile pairs u = u0, u0 + ustep, ... (pair = two column tiles of one 8-row strip)
__device__ __forceinline__ void blk_update(double* C, const double* Xi, const double* Xk, int u0, int ustep, int g, int tg) {
  for (int u = u0; u < 8; u += ustep) {
    const int ms = u >> 1, ns0 = (u & 1) * 2;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    mma_pair_nt(c, Xi + 8 * ms * LDB, LDB, Xk + 8 * ns0 * LDB, Xk + 8 * (ns0 + 1) * LDB, LDB, 0, HB, g, tg);
    double2* p0 = reinterpret_cast<double2*>(C + (8 * ms + g) * LDB + 8 * ns0 + 2 * tg);
    double2 v0 = p0[0], v1 = p0[4];
    v0.x -= c[0]; v0.y -= c[1]; v1.x -= c[2]; v1.y -= c[3];
    p0[0] = v0; p0[4] = v1;
  }
}
// X (32x32 block) = G W^T (W lower triangular, zeros stored above the diagonal): the 8 tile pairs u = u0, u0 + ustep, ...
__device__ __forceinline__ void blk_solve(double* X, const double* G, const double* W, int u0, int ustep, int g, int tg) {
  for (int u = u0; u < 8; u += ustep) {
    const int ms = u >> 1, ns0 = (u & 1) * 2;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    mma_pair_nt(c, G + 8 * ms * LDB, LDB, W + 8 * ns0 * LDB, W + 8 * (ns0 + 1) * LDB, LDB, 0, 8 * (ns0 + 2), g, tg);
    double2* p0 = reinterpret_cast<double2*>(X + (8 * ms + g) * LDB + 8 * ns0 + 2 * tg);
    p0[0] = make_double2(c[0], c[1]); p0[4] = make_double2(c[2], c[3]);
  }
}

// One node = nt tiles (1 or 2) = nbk = 2 nt blocks of 32 columns. Per block column h >= 1, on the critical path only
//   X_h,h-1 = G_h,h-1 W_{h-1}^T ;  G_hh -= X_h,h-1 X_h,h-1^T ;  warp 0: potrf32_sym(G_hh) -> W_h
// while the six warps that do not share warp 0's scheduler (1,2,3,5,6,7; warp 4 would take FP64 issue slots from the chain)
// solve the other blocks of column h-1, apply the rest of its rank-32 update and, once tile a is complete, publish L_ba and
// L_aa^-1 (so that the row tiles of column a and the updates they feed run on other SMs while this CTA factors tile b).
// The first diagonal block is loaded on its own so that the chain starts while the rest of the node is still in flight.
constexpr int F_DW = 6;                 // deferred-work warps
constexpr int F_DT = 32 * F_DW;
__device__ void f_task(const FusedArgs& a, const int* tk, double* smem, unsigned long long* stamps) {
#define F_STAMP(k) do { if (stamps && warp == 0) stamps[k] = (unsigned long long)clock64(); } while (0)   // warp-uniform: no divergence in front of the shuffles
  const int nt = tk[FK_F_NT], nbk = 2 * nt, nblk = nbk * (nbk + 1) / 2;
  double* sG = smem;
  double* sW = sG + nblk * BS;
  double* sX0 = sW + nbk * BS;
  double* sX1 = sX0 + (nbk - 1) * BS;
  auto Xs = [&](int h, int i) { return ((h & 1) ? sX1 : sX0) + (i - h - 1) * BS; };
  __shared__ double sdv[4 * HB];   // 1 / sqrt(d) of every eliminated column of the node
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int dw = warp < 4 ? warp - 1 : warp - 2;          // 0..5 for warps 1,2,3,5,6,7 (meaningless for warps 0 and 4)
  const int dt = 32 * dw + lane;
  const int ta = tk[FK_F_TILE];
  const size_t ld = (size_t)a.ld;
  const double* An = a.A + (size_t)ta * NB * ld + (size_t)ta * NB;   // the node's diagonal super-tile
  F_STAMP(0);
  // ---- load: warp 0 fetches block (0,0) on its own (mirrored into both triangles: potrf32_sym reads full symmetric rows)
  // and starts the chain without waiting for anybody; warps 1..7 bring in the rest of the node behind it ----
  if (warp == 0) {
    double2 val[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) { const int e = lane + 32 * it, rr = e >> 4, cc = (e & 15) * 2; val[it] = ldcg2(An + (size_t)rr * ld + cc); }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int e = lane + 32 * it, rr = e >> 4, cc = (e & 15) * 2;
      if (cc <= rr) { sG[rr * LDB + cc] = val[it].x; sG[cc * LDB + rr] = val[it].x; }
      if (cc + 1 <= rr) { sG[rr * LDB + cc + 1] = val[it].y; sG[(cc + 1) * LDB + rr] = val[it].y; }
    }
    __syncwarp();
  } else {
    for (int e = 512 + tid - 32; e < nblk * 512; e += FTH - 32) {
      const int b = e >> 9, w = e & 511, rr = w >> 4, cc = (w & 15) * 2;
      const int bi = b >= 6 ? 3 : (b >= 3 ? 2 : (b >= 1 ? 1 : 0)), bj = b - bi * (bi + 1) / 2;
      cp_async16(sG + b * BS + rr * LDB + cc, An + (size_t)(HB * bi + rr) * ld + HB * bj + cc);
    }
  }
  F_STAMP(1);
#pragma unroll 1
  for (int h = 0; h < nbk; ++h) {
    if (h > 0) {
      potrf32_finalize(sW + (h - 1) * BS, sdv + HB * (h - 1), tid, FTH);   // raw eliminated rows -> L^-1, by everybody
      __syncthreads();
      blk_solve(Xs(h - 1, h), f_blk(sG, h, h - 1), sW + (h - 1) * BS, warp, 8, g, tg);
      __syncthreads();
      blk_update(f_blk(sG, h, h), Xs(h - 1, h), Xs(h - 1, h), warp, 8, g, tg);   // all 16 tiles: potrf32_sym wants both triangles
      __syncthreads();
    }
    F_STAMP(2 + 2 * h);
    if (warp == 0) {
      potrf32_sym_t<true, false>(f_blk(sG, h, h), sW + h * BS, a.fail, sdv + HB * h);
      F_STAMP(3 + 2 * h);
    } else if (h == 0) {
      // the rest of the node arrives; the other diagonal blocks are mirrored
      cp_async_wait_all();
      bar_named(3, FTH - 32);
      for (int e = tid - 32; e < (nbk - 1) * HB * HB; e += FTH - 32) {
        const int bi = 1 + (e >> 10), rr = (e >> 5) & 31, cc = e & 31;
        if (cc < rr) { double* dg = f_blk(sG, bi, bi); dg[cc * LDB + rr] = dg[rr * LDB + cc]; }
      }
    } else if (warp == 4) {
      if (h == 2) {   // release of L_ba and L_aa^-1 (written by the deferred warps) — off everybody's critical path
        bar_named(4, F_DT + 32);
        if (lane == 0) {
          __threadfence();
          atomicAdd(a.sync + tk[FK_F_XBA], 64);
          atomicAdd(a.sync + tk[FK_F_FIN], 1);
        }
      }
    } else {
      // the other blocks of column h-1 (pair u of block i belongs to deferred warp (8 (i - h - 1) + u) % 6)
      for (int i = h + 1; i < nbk; ++i)
        blk_solve(Xs(h - 1, i), f_blk(sG, i, h - 1), sW + (h - 1) * BS, (dw + F_DW - ((8 * (i - h - 1)) % F_DW)) % F_DW, F_DW, g, tg);
      bar_named(1, F_DT);
      if (h == 2) {   // nt == 2: tile a is complete
        double* Lba = a.A + (size_t)(ta + 1) * NB * ld + (size_t)ta * NB;
        for (int e = dt; e < 4 * 512; e += F_DT) {
          const int b = e >> 9, w = e & 511, rr = w >> 4, cc = (w & 15) * 2;
          const int hh = b & 1, i = 2 + (b >> 1);
          *reinterpret_cast<double2*>(Lba + (size_t)(HB * (i - 2) + rr) * ld + HB * hh + cc) = *reinterpret_cast<const double2*>(Xs(hh, i) + rr * LDB + cc);
        }
        linv_tile(sW, sW + BS, Xs(0, 1), f_blk(sG, 0, 0), a.Linv + (size_t)ta * NB * NB, dw, F_DW, dt, 1);
        bar_arrive(4, F_DT + 32);        // the stores are issued; the idle warp 4 waits for them and raises the flags
      }
      // rest of the rank-32 update of column h-1: blocks (i, k), i >= k >= h, (i, k) != (h, h)
      int base = 0;
      for (int k = h; k < nbk; ++k)
        for (int i = (k == h ? h + 1 : k); i < nbk; ++i) {
          blk_update(f_blk(sG, i, k), Xs(h - 1, i), Xs(h - 1, k), (dw + F_DW - (base % F_DW)) % F_DW, F_DW, g, tg);
          base += 8;
        }
    }
    __syncthreads();
  }
  // ---- last tile of the node: X_qp = G_qp W_p^T of its two blocks is in the column buffer, its inverse goes out ----
  {
    const int p = nbk - 2, q = nbk - 1;
    potrf32_finalize(sW + q * BS, sdv + HB * q, tid, FTH);
    __syncthreads();
    linv_tile(sW + p * BS, sW + q * BS, Xs(p, q), f_blk(sG, p, p), a.Linv + (size_t)(ta + nt - 1) * NB * NB, warp, 8, tid, 2);
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicAdd(a.sync + tk[FK_F_FIN] + nt - 1, 1); }
  }
  F_STAMP(10);
#undef F_STAMP
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
  for (int e = tid; e < nrows * NB / 2; e += FTH) {
    const int rr = e >> 5, cc = (e & 31) * 2;
    cp_async16(sB + rr * LDW + cc, Ag + (size_t)rr * ld + cc);
  }
  for (int e = tid; e < NB * NB / 2; e += FTH) {
    const int rr = e >> 5, cc = (e & 31) * 2;
    cp_async16(sL + rr * LDW + cc, Lg + rr * NB + cc);
  }
  cp_async_wait_all();
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
// waiting / signalling
// ---------------------------------------------------------------------------------------------------------------------
// Every thread takes some of the (sync index, minimum) pairs [d0, d1); returns false when the launch was aborted.
// The first F_INL pairs of the task came with its descriptor (inl[], pair inl0 + k = inl[k]): reading a pair from global memory
// first and polling its counter afterwards is two dependent L2 round trips (~1.5 us) per wait, and an update task waits three times.
constexpr int F_INL = 16;
__device__ bool wait_deps(const FusedArgs& a, int d0, int d1, int* s_abort, const int2* inl, int inl0) {
  for (int e = d0 + (int)threadIdx.x; e < d1; e += FTH) {
    const int2 d = (e - inl0 < F_INL) ? inl[e - inl0] : __ldg(a.deps + e);
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
// U task: 32x32 quadrant of A_ik -= sum_j X_ij X_kj^T
// ---------------------------------------------------------------------------------------------------------------------
constexpr int U_CHUNK = 4;   // source tiles staged at once by a quadrant task (2 x 32 x LDW doubles each)
constexpr int UT_CHUNK = 2;  // ... by a whole-tile task (2 x 64 x LDW doubles each)
// The sources come in two phases (bit 30 of the source = second tile of a pair node, final only when that node's F task ends):
// a chunk never mixes phases, and the task waits for a chunk's row solves right before it stages them, so the first tiles'
// contributions are summed while the second tiles are still being solved. The earlier updates of the target are waited for
// last, before the one read-modify-write of the target.
__device__ bool u_task(const FusedArgs& a, const int* tk, double* smem, int* s_abort, const int2* inl, const int* idx) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int i = tk[FK_U_I], k = tk[FK_U_K], q = tk[FK_U_Q];
  const int e0 = tk[FK_U_SRC0], e1 = tk[FK_U_SRC1];
  const size_t ld = (size_t)a.ld;
  const int nper = (i == k) ? 1 : 2;
  int dep = tk[FK_DEP0];
  // source list: the first F_INL entries came with the descriptor (a read from a.srcs is an L2 round trip, and the chunking loop
  // below chains several of them)
  auto SRC = [&](int e) { return e - e0 < F_INL ? idx[e - e0] : a.srcs[e]; };
  if (q == 4) {
    // ---- whole 64x64 tile: warp tile 32 rows x 16 columns (8 MMA chains per warp) ----
    const int wr = (warp >> 2) * 32, wc = (warp & 3) * 16;
    double acc[4][2][2];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
    for (int c0 = e0; c0 < e1;) {
      int nc = 1;
      while (nc < UT_CHUNK && c0 + nc < e1 && ((SRC(c0 + nc) ^ SRC(c0)) & (1 << 30)) == 0) ++nc;
      if (!wait_deps(a, dep, dep + nc * nper, s_abort, inl, tk[FK_DEP0])) return false;   // (also: the previous chunk has been consumed)
      dep += nc * nper;
      for (int sc = 0; sc < nc; ++sc) {
        const int j = SRC(c0 + sc) & 0x3fffffff;
        double* sA = smem + sc * 2 * NB * LDW;
        double* sB = sA + NB * LDW;
        const double* Xi = a.A + (size_t)i * NB * ld + (size_t)j * NB;
        const double* Xk = a.A + (size_t)k * NB * ld + (size_t)j * NB;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = tid + FTH * u, rr = e >> 5, cc = (e & 31) * 2;
          cp_async16(sA + rr * LDW + cc, Xi + (size_t)rr * ld + cc);
          cp_async16(sB + rr * LDW + cc, Xk + (size_t)rr * ld + cc);
        }
      }
      cp_async_wait_all();
      __syncthreads();
      for (int sc = 0; sc < nc; ++sc) {
        const double* sA = smem + sc * 2 * NB * LDW;
        const double* sB = sA + NB * LDW;
#pragma unroll 2
        for (int k0 = 0; k0 < NB; k0 += 4) {
          double fa[4], fb[2];
#pragma unroll
          for (int x = 0; x < 4; ++x) fa[x] = sA[(wr + 8 * x + g) * LDW + k0 + tg];
#pragma unroll
          for (int y = 0; y < 2; ++y) fb[y] = sB[(wc + 8 * y + g) * LDW + k0 + tg];
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 2; ++y) dmma_m8n8k4(acc[x][y][0], acc[x][y][1], fa[x], fb[y]);
        }
      }
      c0 += nc;
    }
    if (!wait_deps(a, dep, tk[FK_DEP1], s_abort, inl, tk[FK_DEP0])) return false;
    double* C = a.A + (size_t)i * NB * ld + (size_t)k * NB;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        double2* p = reinterpret_cast<double2*>(C + (size_t)(wr + 8 * x + g) * ld + wc + 8 * y + 2 * tg);
        double2 v = __ldcg(p);
        v.x -= acc[x][y][0]; v.y -= acc[x][y][1];
        *p = v;
      }
    return true;
  }
  // ---- one 32x32 quadrant ----
  const int qi = q >> 1, qk = q & 1;
  const int ms = warp >> 1, ns0 = (warp & 1) * 2;
  const bool active = !(i == a.Tn && ms > 0);   // b row: only row 0 (first strip) carries data
  const bool same = (i == k && qi == qk);       // diagonal quadrant of a diagonal tile: both operands are the same rows
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int c0 = e0; c0 < e1;) {
    int nc = 1;
    while (nc < U_CHUNK && c0 + nc < e1 && ((SRC(c0 + nc) ^ SRC(c0)) & (1 << 30)) == 0) ++nc;
    if (!wait_deps(a, dep, dep + nc * nper, s_abort, inl, tk[FK_DEP0])) return false;   // (also: the previous chunk has been consumed)
    dep += nc * nper;
    for (int sc = 0; sc < nc; ++sc) {
      const int j = SRC(c0 + sc) & 0x3fffffff;
      double* sA = smem + sc * 2 * HB * LDW;
      double* sB = sA + HB * LDW;
      const double* Xi = a.A + ((size_t)i * NB + HB * qi) * ld + (size_t)j * NB;
      const double* Xk = a.A + ((size_t)k * NB + HB * qk) * ld + (size_t)j * NB;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = tid + FTH * u, rr = e >> 5, cc = (e & 31) * 2;
        cp_async16(sA + rr * LDW + cc, Xi + (size_t)rr * ld + cc);
        if (!same) cp_async16(sB + rr * LDW + cc, Xk + (size_t)rr * ld + cc);
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (active) {
      for (int sc = 0; sc < nc; ++sc) {
        const double* sA = smem + sc * 2 * HB * LDW;
        const double* sB = same ? sA : sA + HB * LDW;
        mma_pair_nt(acc, sA + 8 * ms * LDW, LDW, sB + 8 * ns0 * LDW, sB + 8 * (ns0 + 1) * LDW, LDW, 0, NB, g, tg);
      }
    }
    c0 += nc;
  }
  if (!wait_deps(a, dep, tk[FK_DEP1], s_abort, inl, tk[FK_DEP0])) return false;
  if (active) {
    double* C = a.A + ((size_t)i * NB + HB * qi) * ld + (size_t)k * NB + HB * qk;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      double2* p = reinterpret_cast<double2*>(C + (size_t)(8 * ms + g) * ld + 8 * (ns0 + t) + 2 * tg);
      double2 v = __ldcg(p);
      v.x -= acc[2 * t]; v.y -= acc[2 * t + 1];
      *p = v;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// B task: x_j = L_jj^-T (y_j - sum_{i in below(j)} L_ij^T x_i)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int B_PRE = 4;   // L tiles staged in shared memory while the task waits for the x_i (the rest is read from L2 afterwards)
__device__ bool b_task(const FusedArgs& a, const int* tk, double* smem, int* s_abort, const int2* inl, const int* idx) {
  __shared__ double sx[4][NB];
  __shared__ double st[4][NB];
  __shared__ double stt[NB];
  const int tid = threadIdx.x, c = tid & 63, g = tid >> 6;
  const int j = tk[FK_B_J], e0 = tk[FK_B_BEL0], e1 = tk[FK_B_BEL1];
  const size_t ld = (size_t)a.ld;
  double* sM = smem;                    // L_jj^-1, tight
  double* sT = smem + NB * NB;          // up to B_PRE tiles, tight
  // inputs that are final before any x_i is: L_jj^-1, the L_ij tiles, y_j (the first two dependencies + the xdone of the tiles)
  auto BEL = [&](int e) { return e - e0 < F_INL ? idx[e - e0] : a.below[e]; };   // row tiles below j: first F_INL with the descriptor
  const int dmid = tk[FK_DEP0] + 1 + (e1 - e0) + 1;   // [fin_j, xdone(Tn, j), xdone(i, j)...] then [bx_i...]
  if (!wait_deps(a, tk[FK_DEP0], dmid, s_abort, inl, tk[FK_DEP0])) return false;
  {
    const double* M = a.Linv + (size_t)j * NB * NB;
    for (int e = tid; e < NB * NB / 2; e += FTH) cp_async16(sM + 2 * e, M + 2 * e);
    const int npre = min(e1 - e0, B_PRE);
    for (int t = 0; t < npre; ++t) {
      const double* L = a.A + (size_t)BEL(e0 + t) * NB * ld + (size_t)j * NB;
      for (int e = tid; e < NB * NB / 2; e += FTH) {
        const int rr = e >> 5, cc = (e & 31) * 2;
        cp_async16(sT + t * NB * NB + rr * NB + cc, L + (size_t)rr * ld + cc);
      }
    }
    cp_async_wait_all();
  }
  const double yj = (g == 0) ? __ldcg(a.A + (size_t)a.Tn * NB * ld + (size_t)j * NB + c) : 0.0;   // y_j = row 0 of the solved b tile
  // x of the other nodes first; the partner tile of a pair (below[e0], the tile this node's F task factored last) comes on its
  // own after them, so that everything but its 64x64 term is already summed when it arrives
  const int npart = tk[FK_B_PARTNER];
  if (!wait_deps(a, dmid, tk[FK_DEP1] - npart, s_abort, inl, tk[FK_DEP0])) return false;
  double t0 = 0.0, t1 = 0.0;
  for (int e = e0 + npart + g; e < e1; e += 4) {
    const int i = BEL(e);
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
  if (npart) {
    if (!wait_deps(a, tk[FK_DEP1] - 1, tk[FK_DEP1], s_abort, inl, tk[FK_DEP0])) return false;
    // 64 x 64 term of the partner, rows split over the four groups (its tile is the first staged one)
    sx[g][c] = __ldcg(a.x + (size_t)BEL(e0) * NB + c);
    bar_named(1 + g, 64);
    const double* L = sT + c;
#pragma unroll
    for (int r = 16 * g; r < 16 * g + 16; r += 2) { t0 -= L[r * NB] * sx[g][r]; t1 -= L[(r + 1) * NB] * sx[g][r + 1]; }
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
  __shared__ int2 s_inl[F_INL];   // the task's first dependency pairs, fetched with the descriptor
  __shared__ int s_idx[F_INL];    // ... and its first source tiles (U) / row tiles below (B)
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
    else if (tid < F_TASK_INTS + F_INL) s_inl[tid - F_TASK_INTS] = __ldg(a.dep_inl + (size_t)t * F_INL + tid - F_TASK_INTS);
    else if (tid < F_TASK_INTS + 2 * F_INL) s_idx[tid - F_TASK_INTS - F_INL] = __ldg(a.idx_inl + (size_t)t * F_INL + tid - F_TASK_INTS - F_INL);
    __syncthreads();
    const int type = s_task[FK_TYPE];
    unsigned long long t_pop = 0, t_ready = 0;
    if (a.trace && tid == 0) t_pop = gtime();
    if (type == FT_B) {
      if (!b_task(a, s_task, smem, &s_abort, s_inl, s_idx)) break;
    } else if (type == FT_U) {
      if (!u_task(a, s_task, smem, &s_abort, s_inl, s_idx)) break;
    } else {
      if (!wait_deps(a, s_task[FK_DEP0], s_task[FK_DEP1], &s_abort, s_inl, s_task[FK_DEP0])) break;
      if (a.trace && tid == 0) t_ready = gtime();
      if (type == FT_S) s_task_run(a, s_task, smem);
      else f_task(a, s_task, smem, a.trace ? a.trace + TRACE_WORDS * (size_t)t + 4 : nullptr);
    }
    __syncthreads();   // all global stores of the task are issued
    if (tid == 0) {
      if (s_task[FK_SIG] >= 0) {
        __threadfence();
        const int nsig = (type == FT_U && s_task[FK_U_Q] == 4) ? 4 : 1;   // a whole-tile update bumps the counters of its four quadrants
        for (int x = 0; x < nsig; ++x) red_release_add_s32(a.sync + s_task[FK_SIG] + x, s_task[FK_SIGINC]);
      }
      if (a.trace) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        unsigned long long* tr = a.trace + TRACE_WORDS * (size_t)t;
        tr[0] = t_pop; tr[1] = t_ready; tr[2] = gtime(); tr[3] = smid;
      }
    }
  }
}

constexpr int FUSED_SMEM_DOUBLES = f_layout_doubles(4) > (1 + B_PRE) * NB * NB ? f_layout_doubles(4) : (1 + B_PRE) * NB * NB;
static_assert(FUSED_SMEM_DOUBLES * 8 <= 227 * 1024, "fused Cholesky exceeds the shared memory of one SM");
static_assert(U_CHUNK * 2 * HB * LDW <= FUSED_SMEM_DOUBLES && UT_CHUNK * 2 * NB * LDW <= FUSED_SMEM_DOUBLES && 2 * NB * LDW <= FUSED_SMEM_DOUBLES, "S / U staging must fit");

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
  {   // first F_INL dependency pairs of every task in queue order (padding: the queue head, which is never below 0)
    std::vector<int2> inl((size_t)H.f_ntasks * F_INL, make_int2(FS_HEAD, 0));
    for (int t = 0; t < H.f_ntasks; ++t) {
      const int d0 = H.f_tasks[(size_t)t * F_TASK_INTS + FK_DEP0], d1 = H.f_tasks[(size_t)t * F_TASK_INTS + FK_DEP1];
      for (int e = d0; e < d1 && e - d0 < F_INL; ++e) inl[(size_t)t * F_INL + e - d0] = make_int2(H.f_deps[e].x, H.f_deps[e].y);
    }
    TSL_CUDA(sym->f_dep_inl.upload(inl.data(), inl.size(), s));
    std::vector<int> ix((size_t)H.f_ntasks * F_INL, 0);
    for (int t = 0; t < H.f_ntasks; ++t) {
      const int* rec = &H.f_tasks[(size_t)t * F_TASK_INTS];
      if (rec[FK_TYPE] == FT_U) for (int e = rec[FK_U_SRC0]; e < rec[FK_U_SRC1] && e - rec[FK_U_SRC0] < F_INL; ++e) ix[(size_t)t * F_INL + e - rec[FK_U_SRC0]] = H.f_srcs[e];
      if (rec[FK_TYPE] == FT_B) for (int e = rec[FK_B_BEL0]; e < rec[FK_B_BEL1] && e - rec[FK_B_BEL0] < F_INL; ++e) ix[(size_t)t * F_INL + e - rec[FK_B_BEL0]] = H.f_below[e];
    }
    TSL_CUDA(sym->f_idx_inl.upload(ix.data(), ix.size(), s));
    TSL_CUDA(cudaStreamSynchronize(s));   // `inl` is a local
  }
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
  a.A = A; a.ld = ld; a.Tn = Tn; a.tasks = sym.f_tasks.p; a.ntasks = sym.f_ntasks; a.deps = sym.f_deps.p; a.dep_inl = sym.f_dep_inl.p; a.idx_inl = sym.f_idx_inl.p; a.srcs = sym.f_srcs.p; a.below = sym.f_below.p;
  a.sync = sym.f_sync.p; a.Linv = sym.Ldiag.p; a.x = xout; a.fail = d_fail; a.trace = trace;
  const int grid = std::max(1, std::min(ctx->sm_count, sym.f_ntasks));
  LAUNCH(launch_k(chol_fused_kernel, grid, FTH, smem, ctx->stream, a));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl
