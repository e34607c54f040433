// Micro-benchmark of the 64x64 tile routines behind potrf_trsm_kernel (textslam_b200/csrc/chol_tile.cuh): per-phase
// clock64() stamps of one CTA + a numerical check against a host Cholesky. Variants are compared on the same input.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o chol_tile_bench chol_tile_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../textslam_b200/csrc/chol_tile.cuh"
using namespace tsl;

#define STAMP(k) do { if (threadIdx.x == 0) stamps[(k)] = clock64(); } while (0)

template <int V>
__global__ void __launch_bounds__(PT_THREADS) bench_kernel(const double* __restrict__ Ajj, const double* __restrict__ Aij, double* __restrict__ Lout,
                                                           double* __restrict__ Xout, long long* __restrict__ stamps, int* fail) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double sinv[NB];
  const int ld = NB;
  stamps += blockIdx.x * 16;
  STAMP(0);
  if (V == 0) {   // first generation: Crout potrf32 with block barriers, left-looking trsm32, 4x4 gemm (LDT = 65)
    double* sT = smem;
    double* sX = smem + NB * LDT;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double vt[16], vx[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int e = threadIdx.x + PT_THREADS * (16 * half + u), r = e >> 6, c = e & 63;
        vt[u] = (c <= r) ? Ajj[(size_t)r * ld + c] : 0.0;
        vx[u] = Aij[(size_t)r * ld + c];
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int e = threadIdx.x + PT_THREADS * (16 * half + u), r = e >> 6, c = e & 63;
        sT[r * LDT + c] = vt[u]; sX[r * LDT + c] = vx[u];
      }
    }
    __syncthreads();
    STAMP(1);
    potrf32(sT, sinv, fail);
    STAMP(2);
    trsm32(sT + HB * LDT, HB, sT, sinv);
    __syncthreads();
    STAMP(3);
    gemm_nt32(sT + HB * LDT + HB, sT + HB * LDT, sT + HB * LDT, HB);
    __syncthreads();
    STAMP(4);
    potrf32(sT + HB * LDT + HB, sinv + HB, fail);
    __syncthreads();
    trsm32(sX, NB, sT, sinv);
    __syncthreads();
    gemm_nt32(sX + HB, sX, sT + HB * LDT, NB);
    __syncthreads();
    STAMP(5);
    trsm32(sX + HB, NB, sT + HB * LDT + HB, sinv + HB);
    __syncthreads();
    STAMP(6);
    for (int e = threadIdx.x; e < NB * NB; e += PT_THREADS) { const int r = e >> 6, c = e & 63; Xout[(size_t)r * ld + c] = sX[r * LDT + c]; Lout[(size_t)r * ld + c] = sT[r * LDT + c]; }
  } else {        // second generation (factor_solve_tile)
    double* sT = smem;
    double* sX = smem + NB * LD2;
    double* sLt = smem + 2 * NB * LD2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double2 vt[8], vx[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = threadIdx.x + PT_THREADS * (8 * half + u), r = e >> 5, c = (e & 31) * 2;
        vt[u] = *reinterpret_cast<const double2*>(Ajj + (size_t)r * ld + c);
        vx[u] = *reinterpret_cast<const double2*>(Aij + (size_t)r * ld + c);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = threadIdx.x + PT_THREADS * (8 * half + u), r = e >> 5, c = (e & 31) * 2;
        *reinterpret_cast<double2*>(sT + r * LD2 + c) = make_double2(c <= r ? vt[u].x : 0.0, c + 1 <= r ? vt[u].y : 0.0);
        *reinterpret_cast<double2*>(sX + r * LD2 + c) = vx[u];
      }
    }
    __syncthreads();
    STAMP(1);
    factor_solve_tile<true>(sT, sX, sLt, sinv, fail, stamps);
    for (int e = threadIdx.x; e < NB * NB; e += PT_THREADS) { const int r = e >> 6, c = e & 63; Xout[(size_t)r * ld + c] = sX[r * LD2 + c]; Lout[(size_t)r * ld + c] = sT[r * LD2 + c]; }
  }
  __syncthreads();
  STAMP(7);
}

__global__ void rsqrt_check_kernel(const double* d, double* out_fast, double* out_lib, long long* cyc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { out_fast[i] = rsqrt_pivot(d[i]); out_lib[i] = rsqrt(d[i]); }
  if (i == 0) {   // dependent-chain latency of both
    double x = d[0];
    long long t0 = clock64();
#pragma unroll 1
    for (int k = 0; k < 64; ++k) x = rsqrt_pivot(x) + 1.5;
    long long t1 = clock64();
    double y = d[0];
#pragma unroll 1
    for (int k = 0; k < 64; ++k) y = rsqrt(y) + 1.5;
    long long t2 = clock64();
    cyc[0] = (t1 - t0) / 64; cyc[1] = (t2 - t1) / 64; out_fast[n] = x + y;
  }
}

template <int V>
static void run(const char* name, int grid, const double* dA, const double* dB, double* dL, double* dX, long long* dS, int* dF,
                const std::vector<double>& Lref, const std::vector<double>& Xref) {
  const int smem = 3 * NB * LD2 * sizeof(double);
  cudaFuncSetAttribute(bench_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  std::vector<long long> st(16 * grid);
  for (int rep = 0; rep < 6; ++rep) {
    cudaMemset(dL, 0, NB * NB * 8); cudaMemset(dX, 0, NB * NB * 8);
    cudaEventRecord(e0);
    bench_kernel<V><<<grid, PT_THREADS, smem>>>(dA, dB, dL, dX, dS, dF);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); return; }
  cudaMemcpy(st.data(), dS, st.size() * 8, cudaMemcpyDeviceToHost);
  std::vector<double> L(NB * NB), X(NB * NB);
  cudaMemcpy(L.data(), dL, NB * NB * 8, cudaMemcpyDeviceToHost); cudaMemcpy(X.data(), dX, NB * NB * 8, cudaMemcpyDeviceToHost);
  double eL = 0, eX = 0;
  for (int r = 0; r < NB; ++r) for (int c = 0; c <= r; ++c) eL = std::max(eL, std::fabs(L[r * NB + c] - Lref[r * NB + c]) / (1e-300 + std::fabs(Lref[r * NB + c]) + 1e-3));
  for (int i = 0; i < NB * NB; ++i) eX = std::max(eX, std::fabs(X[i] - Xref[i]) / (std::fabs(Xref[i]) + 1e-3));
  static const char* ph0[7] = {"load", "potrf32#1", "trsm(L21)", "gemm(A22)", "potrf32#2+trsm(X1)+gemm(X2)", "trsm(X2)", "store"};
  static const char* ph1[7] = {"load", "potrf32#1", "trsm(L21)|trsm(X1)", "gemm(A22)+gemm(X2)", "potrf32#2", "trsm(X2)", "store"};
  printf("%-28s grid %3d  kernel %.2f us  errL %.1e errX %.1e  | cycles:", name, grid, best * 1e3, eL, eX);
  for (int k = 0; k < 7; ++k) printf(" %s %lld", (V == 0 ? ph0 : ph1)[k], st[k + 1] - st[k]);
  printf(" | total %lld\n", st[7] - st[0]);
}

int main() {
  std::vector<double> G(NB * NB), A(NB * NB), B(NB * NB), L(NB * NB, 0.0), X(NB * NB);
  srand(1);
  for (auto& v : G) v = rand() / (double)RAND_MAX - 0.5;
  for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < NB; ++i) for (int j = 0; j < NB; ++j) { double s = (i == j) ? 4.0 : 0.0; for (int k = 0; k < NB; ++k) s += G[i * NB + k] * G[j * NB + k]; A[i * NB + j] = s; }
  for (int j = 0; j < NB; ++j) {
    double d = A[j * NB + j]; for (int k = 0; k < j; ++k) d -= L[j * NB + k] * L[j * NB + k];
    L[j * NB + j] = std::sqrt(d);
    for (int i = j + 1; i < NB; ++i) { double s = A[i * NB + j]; for (int k = 0; k < j; ++k) s -= L[i * NB + k] * L[j * NB + k]; L[i * NB + j] = s / L[j * NB + j]; }
  }
  for (int r = 0; r < NB; ++r) for (int c = 0; c < NB; ++c) { double s = B[r * NB + c]; for (int k = 0; k < c; ++k) s -= X[r * NB + k] * L[c * NB + k]; X[r * NB + c] = s / L[c * NB + c]; }
  double *dA, *dB, *dL, *dX; long long* dS; int* dF;
  cudaMalloc(&dA, NB * NB * 8); cudaMalloc(&dB, NB * NB * 8); cudaMalloc(&dL, NB * NB * 8); cudaMalloc(&dX, NB * NB * 8); cudaMalloc(&dS, 16 * 256 * 8); cudaMalloc(&dF, 4);
  cudaMemcpy(dA, A.data(), NB * NB * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), NB * NB * 8, cudaMemcpyHostToDevice); cudaMemset(dF, 0, 4);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock %d kHz\n", clk);
  for (int grid : {1, 40}) {
    run<0>("gen1 (crout, block barriers)", grid, dA, dB, dL, dX, dS, dF, L, X);
    run<1>("gen2 (factor_solve_tile)", grid, dA, dB, dL, dX, dS, dF, L, X);
  }
  {
    const int n = 1 << 16;
    std::vector<double> d(n), f1(n + 1), f2(n);
    for (int i = 0; i < n; ++i) d[i] = std::exp((rand() / (double)RAND_MAX - 0.5) * 60.0);
    double *dd, *o1, *o2; long long* cy;
    cudaMalloc(&dd, n * 8); cudaMalloc(&o1, (n + 1) * 8); cudaMalloc(&o2, n * 8); cudaMalloc(&cy, 16);
    cudaMemcpy(dd, d.data(), n * 8, cudaMemcpyHostToDevice);
    rsqrt_check_kernel<<<n / 256, 256>>>(dd, o1, o2, cy, n);
    cudaMemcpy(f1.data(), o1, n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(f2.data(), o2, n * 8, cudaMemcpyDeviceToHost);
    long long c[2]; cudaMemcpy(c, cy, 16, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int i = 0; i < n; ++i) {
      const long double ref = 1.0L / sqrtl((long double)d[i]);
      const double ulp = std::ldexp(1.0, std::ilogb((double)ref) - 52);
      e1 = std::max(e1, (double)fabsl((long double)f1[i] - ref) / ulp); e2 = std::max(e2, (double)fabsl((long double)f2[i] - ref) / ulp);
    }
    printf("rsqrt: rsqrt_pivot max err %.2f ulp, %lld cycles/dependent call; library rsqrt max err %.2f ulp, %lld cycles\n", e1, c[0], e2, c[1]);
  }
  int f = 0; cudaMemcpy(&f, dF, 4, cudaMemcpyDeviceToHost); printf("fail flag %d\n", f);
  return 0;
}
