// Micro-benchmark behind potrf32_sym (textslam_b200/csrc/chol_potrf.cuh): one warp, clock64() around the routine and around
// its ingredients (loop-carried chain, shuffle, reciprocal, single-warp DFMA and LDS.128 issue rates), plus a numerical check of
// the factor inverse against a host Cholesky.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o potrf_sym_bench potrf_sym_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../textslam_b200/csrc/chol_potrf.cuh"
using namespace tsl;

template <int VAR>
__global__ void __launch_bounds__(32, 1) k_potrf(const double* Gg, double* Wg, long long* out, int* fail, int reps) {
  __shared__ __align__(16) double G[32 * LDB];
  __shared__ __align__(16) double W[32 * LDB];
  __shared__ double dv[32];
  for (int i = threadIdx.x; i < 32 * LDB; i += 32) G[i] = Gg[i];
  __syncwarp();
  long long best = 1ll << 60;
  for (int it = 0; it < reps; ++it) {
    if (VAR >= 2) { for (int i = threadIdx.x; i < 32 * LDB; i += 32) G[i] = Gg[i]; }   // the blocked variant works in place
    __syncwarp();
    const long long t0 = clock64();
    if (VAR == 2) potrf32_blk(G, W, fail);
    else if (VAR == 3) potrf32_blk_t<9, true, true, true>(G, W, fail);
    else if (VAR == 4) potrf32_blk_t<12, false, true, true>(G, W, fail);
    else if (VAR == 5) potrf32_blk_t<12, false, false, true>(G, W, fail);
    else if (VAR == 6) potrf32_blk_t<12, false, false, false>(G, W, fail);
    else if (VAR == 8) potrf32_sym_t<true, false>(G, W, fail, dv);
    else if (VAR == 9) potrf32_sym_t<false, false>(G, W, fail, dv);
    else potrf32_sym_t<VAR == 0>(G, W, fail);
    __syncwarp();
    const long long t1 = clock64();
    best = min(best, t1 - t0);
  }
  if (VAR >= 8) { __syncwarp(); potrf32_finalize(W, dv, threadIdx.x, 32); __syncwarp(); }
  for (int i = threadIdx.x; i < 32 * LDB; i += 32) Wg[i] = W[i];
  if (threadIdx.x == 0) out[0] = best;
}

// ingredient timings; mode selects the loop body. 32 dependent iterations like the 32 columns of a block.
template <int MODE>
__global__ void __launch_bounds__(32, 1) k_part(double seed, double* sink, long long* out) {
  __shared__ __align__(16) double buf[64];
  const int r = threadIdx.x;
  buf[r] = seed + r; buf[r + 32] = seed - r;
  __syncwarp();
  double d = seed + 2.0, x = seed + 3.0 + r, p1 = 1e-3 * (r + 1);
  double v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = seed + k + r;
  double s = 1e-9 * (r + 1);
  long long best = 1ll << 60;
  for (int rep = 0; rep < 5; ++rep) {
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int base = 0; base < 32; base += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = base + j;
        if (MODE == 0) {          // chain as in potrf32_sym: reciprocal -> fma -> shuffle -> positivity select
          const double rinv = rcp_pivot(d);
          x = fma(-p1, rinv, x);
          double dn = __shfl_sync(0xffffffffu, x, (c + 1) & 31);
          d = dn > 0.0 ? dn : 1.0;
        } else if (MODE == 1) {   // reciprocal + fma only
          const double rinv = rcp_pivot(d);
          d = fma(-p1, rinv, d) + 2.0;
        } else if (MODE == 2) {   // shuffle only (64-bit = two SHFL)
          d = __shfl_sync(0xffffffffu, d, (c + 1) & 31) + 1.0;
        } else if (MODE == 3) {   // 30 independent DFMAs per step, one warp
#pragma unroll
          for (int p = 2; p < 32; ++p) v[p] = fma(-s, d, v[p]);
        } else if (MODE == 4) {   // 15 broadcast LDS.128 + 30 DFMAs per step (the update part of a column)
          const double* cb = buf + base;
          double col[32];
#pragma unroll
          for (int p = 2; p < 32; p += 2) { const double2 t = *reinterpret_cast<const double2*>(cb + p); col[p] = t.x; col[p + 1] = t.y; }
#pragma unroll
          for (int p = 2; p < 32; ++p) v[p] = fma(-s, col[p], v[p]);
          buf[(r + c) & 63] = v[2 + (c & 7)];
          __syncwarp();
        } else if (MODE == 5) {   // raw MUFU.RCP64H dependent chain
          double y0;
          asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
          d = y0;
        } else if (MODE == 6) {   // dependent DFMA chain
          d = fma(d, 1.0000001, 1e-9);
        }
      }
    }
    __syncwarp();
    const long long t1 = clock64();
    best = min(best, t1 - t0);
  }
  double acc = d + x;
#pragma unroll
  for (int k = 0; k < 32; ++k) acc += v[k];
  sink[r] = acc;
  if (r == 0) out[0] = best;
}

// potrf32_sym's loop with parts switched off (timing only, results are not meaningful): F_FMA = update FMAs, F_SM = column publish
// through shared memory (STS, warp barrier, LDS), F_SEL = the selects / pivot bookkeeping
template <bool F_FMA, bool F_SM, bool F_SEL>
__global__ void __launch_bounds__(32, 1) k_ablate(double seed, double* sink, long long* out) {
  __shared__ __align__(16) double colbuf[2][64];
  const int r = threadIdx.x;
  const unsigned full = 0xffffffffu;
  double v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = seed + 0.01 * k + 0.02 * r + (k == r ? 40.0 : 0.0);
  colbuf[0][r] = v[0]; colbuf[0][r + 32] = v[0]; colbuf[1][r] = v[1]; colbuf[1][r + 32] = v[1];
  __syncwarp();
  long long best = 1ll << 60;
  for (int rep = 0; rep < 5; ++rep) {
    double d = 40.0 + seed, rinv = 1.0 / d, vc = 1e-3 * r, p1 = 1e-4 * r, dr = 0.0;
    bool bad = false;
    double colA[32], colB[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) { colA[p] = 1e-3 * p; colB[p] = 2e-3 * p; }
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int base = 0; base < 32; base += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = base + j;
        double* cur = (j & 1) ? colB : colA;
        double* nxt = (j & 1) ? colA : colB;
        v[j + 1] = fma(-p1, rinv, v[j + 1]);
        const double s = vc * rinv;
        double dn = __shfl_sync(full, v[j + 1], (c + 1) & 31);
        const double bn = __shfl_sync(full, v[j + 1], (c + 2) & 31);
        if (F_SM) {
          const double pv = (r > c) ? v[j + 1] : 0.0;
          colbuf[(c + 1) & 1][r] = pv; colbuf[(c + 1) & 1][r + 32] = pv;
          __syncwarp();
          const int jn = (j + 1) & 7;
          const double* cb = &colbuf[(c + 1) & 1][c + 1 - jn];
#pragma unroll
          for (int p = (jn + 2) & ~1; p < 32; p += 2) { const double2 t = *reinterpret_cast<const double2*>(cb + p); nxt[p] = t.x; nxt[p + 1] = t.y; }
        }
        if (F_SEL) {
          bad |= (c + 1 < 32) & !(dn > 0.0);
          dn = dn > 0.0 ? dn : 1.0;
          if (r == c + 1) dr = dn;
        }
        dn = fabs(dn) + 30.0;   // keep the synthetic pivots sane
        const double rinv_n = rcp_pivot(dn);
        if (F_FMA) {
#pragma unroll
          for (int p = j + 2; p < 32; ++p) v[p] = fma(-s, cur[p], v[p]);
        }
        vc = F_SEL ? ((r == c + 1 || c + 2 >= 32) ? 0.0 : v[j + 1]) : v[j + 1];
        p1 = vc * bn * 1e-6;
        rinv = rinv_n;
      }
      double t8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t8[q] = v[q];
#pragma unroll
      for (int p = 0; p < 24; ++p) v[p] = v[p + 8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[24 + q] = t8[q];
    }
    __syncwarp();
    const long long t1 = clock64();
    best = min(best, t1 - t0);
    double acc = dr + (bad ? 1.0 : 0.0);
#pragma unroll
    for (int k = 0; k < 32; ++k) acc += v[k] + colA[k] + colB[k];
    sink[r] = acc;
  }
  if (r == 0) out[0] = best;
}

// two-warp blocked variant: warp 0 = chain, warp 1 = helper
__global__ void __launch_bounds__(64, 1) k_potrf2(const double* Gg, double* Wg, long long* out, int* fail, int reps) {
  __shared__ __align__(16) double G[32 * LDB];
  __shared__ __align__(16) double W[32 * LDB];
  const int warp = threadIdx.x >> 5;
  long long best = 1ll << 60;
  for (int it = 0; it < reps; ++it) {
    for (int i = threadIdx.x; i < 32 * LDB; i += 64) G[i] = Gg[i];
    __syncthreads();
    const long long t0 = clock64();
    potrf32_blk2(G, W, fail, warp, 1);
    __syncwarp();
    const long long t1 = clock64();
    if (warp == 0) best = min(best, t1 - t0);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 32 * LDB; i += 64) Wg[i] = W[i];
  if (threadIdx.x == 0) out[0] = best;
}

int main() {
  const int n = 32;
  std::vector<double> A(n * n), G(32 * LDB, 0.0), W(32 * LDB), L(n * n, 0.0);
  srand(1);
  std::vector<double> B(n * n);
  for (auto& x : B) x = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k];
      A[i * n + j] = s + (i == j ? 1.0 : 0.0);
    }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) G[i * LDB + j] = A[i * n + j];
  for (int j = 0; j < n; ++j) {   // host Cholesky
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    L[j * n + j] = sqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = A[i * n + j];
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / L[j * n + j];
    }
  }
  double *dG, *dW, *dsink; long long* dout; int* dfail;
  cudaMalloc(&dG, G.size() * 8); cudaMalloc(&dW, W.size() * 8); cudaMalloc(&dsink, 32 * 8); cudaMalloc(&dout, 8); cudaMalloc(&dfail, 4);
  cudaMemcpy(dG, G.data(), G.size() * 8, cudaMemcpyHostToDevice); cudaMemset(dfail, 0, 4);
  long long cyc = 0;
  for (int variant = 0; variant < 10; ++variant) {
    switch (variant) {
      case 0: k_potrf<0><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      case 1: k_potrf<1><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      case 2: k_potrf<2><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      case 3: k_potrf<3><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      case 4: k_potrf<4><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      case 5: k_potrf<5><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      case 6: k_potrf<6><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      case 7: k_potrf2<<<1, 64>>>(dG, dW, dout, dfail, 10); break;
      case 8: k_potrf<8><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
      default: k_potrf<9><<<1, 32>>>(dG, dW, dout, dfail, 10); break;
    }
    cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(W.data(), dW, W.size() * 8, cudaMemcpyDeviceToHost);
    double err = 0;   // || W L - I ||_max
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < n; ++k) s += W[i * LDB + k] * L[k * n + j];
        err = fmax(err, fabs(s - (i == j ? 1.0 : 0.0)));
      }
    printf("potrf32_sym (%s): %lld cycles (%.1f per column), ||W L - I||_max = %.2e, %s\n", (const char*[]){"rotating frame", "straight line", "blocked, panels of 8", "blocked, record stride 9", "blocked, no rank-8 update (timing only)", "blocked, no update, no epilogue", "blocked, chain + records only", "blocked, two warps (chain + helper)", "rotating frame, raw rows out (finalised outside the timed part)", "straight line, raw rows out"}[variant], cyc, cyc / 32.0, err,
           cudaGetErrorString(cudaGetLastError()));
  }
  {
    const char* an[5] = {"chain only (fma, 2 shuffles, rcp)", "+ selects", "+ shared-memory column publish", "+ 30 FMAs (no shared memory)", "everything"};
    for (int m = 0; m < 5; ++m) {
      switch (m) {
        case 0: k_ablate<false, false, false><<<1, 32>>>(1.5, dsink, dout); break;
        case 1: k_ablate<false, false, true><<<1, 32>>>(1.5, dsink, dout); break;
        case 2: k_ablate<false, true, true><<<1, 32>>>(1.5, dsink, dout); break;
        case 3: k_ablate<true, false, true><<<1, 32>>>(1.5, dsink, dout); break;
        default: k_ablate<true, true, true><<<1, 32>>>(1.5, dsink, dout); break;
      }
      cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
      printf("  ablation: %-40s %7.1f cycles per column (%s)\n", an[m], cyc / 32.0, cudaGetErrorString(cudaGetLastError()));
    }
  }
  const char* names[7] = {"chain (rcp -> fma -> shfl -> select)", "rcp_pivot + fma", "64-bit shuffle + add", "30 independent DFMA", "15 LDS.128 + 30 DFMA + STS + syncwarp", "MUFU.RCP64H", "dependent DFMA"};
  for (int m = 0; m < 7; ++m) {
    switch (m) {
      case 0: k_part<0><<<1, 32>>>(1.5, dsink, dout); break;
      case 1: k_part<1><<<1, 32>>>(1.5, dsink, dout); break;
      case 2: k_part<2><<<1, 32>>>(1.5, dsink, dout); break;
      case 3: k_part<3><<<1, 32>>>(1.5, dsink, dout); break;
      case 4: k_part<4><<<1, 32>>>(1.5, dsink, dout); break;
      case 5: k_part<5><<<1, 32>>>(1.5, dsink, dout); break;
      default: k_part<6><<<1, 32>>>(1.5, dsink, dout); break;
    }
    cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
    printf("  %-42s %7.1f cycles per step (%s)\n", names[m], cyc / 32.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
