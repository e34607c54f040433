// Micro-benchmark (B200, sm_100a): FP64 FMA vs FP64 tensor (mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16) issue rates
// and dependent-chain latencies. Used to decide how the Cholesky tile update is written.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters) {
  double a[ILP], b = 1.000001, c = 0.999999;
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
  double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dmma884_kernel(double* out, int iters) {
  double d[ILP][2]; double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = 0;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void dmma16816_kernel(double* out, int iters) {
  double d[ILP][4]; double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - (threadIdx.x + i) * 1e-9;
  for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  double s = 0; for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount; double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 4096;
  printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
  for (int warps : {1, 4, 8, 16, 32}) {
    float ms = time_it([&] { dfma_kernel<8><<<sms, warps * 32>>>(out, iters); });
    double flop = 2.0 * 8 * iters * warps * 32 * sms;
    printf("DFMA   ilp8  warps/SM=%2d : %8.3f ms  %8.2f TFLOP/s  (%.1f FMA/clk/SM @1.9GHz)\n", warps, ms, flop / ms / 1e9, flop / 2 / (ms * 1e-3) / sms / 1.9e9);
  }
  { float ms = time_it([&] { dfma_kernel<1><<<sms, 32>>>(out, iters); }); printf("DFMA dependent chain, 1 warp/SM: %.1f cycles/FMA @1.9GHz\n", ms * 1e-3 * 1.9e9 / iters); }
  for (int warps : {1, 4, 8, 16}) {
    float ms = time_it([&] { dmma884_kernel<8><<<sms, warps * 32>>>(out, iters); });
    double flop = 2.0 * 8 * 8 * 4 * 8 * iters * warps * sms;
    printf("DMMA m8n8k4  ilp8 warps/SM=%2d : %8.3f ms  %8.2f TFLOP/s  (%.1f cycles per mma per warp)\n", warps, ms, flop / ms / 1e9, ms * 1e-3 * 1.9e9 / (8.0 * iters));
  }
  { float ms = time_it([&] { dmma884_kernel<1><<<sms, 32>>>(out, iters); }); printf("DMMA m8n8k4 dependent chain, 1 warp/SM: %.1f cycles/mma\n", ms * 1e-3 * 1.9e9 / iters); }
  for (int warps : {1, 4, 8, 16}) {
    float ms = time_it([&] { dmma16816_kernel<4><<<sms, warps * 32>>>(out, iters); });
    double flop = 2.0 * 16 * 8 * 16 * 4 * iters * warps * sms;
    printf("DMMA m16n8k16 ilp4 warps/SM=%2d : %8.3f ms  %8.2f TFLOP/s  (%.1f cycles per mma per warp)\n", warps, ms, flop / ms / 1e9, ms * 1e-3 * 1.9e9 / (4.0 * iters));
  }
  { float ms = time_it([&] { dmma16816_kernel<1><<<sms, 32>>>(out, iters); }); printf("DMMA m16n8k16 dependent chain, 1 warp/SM: %.1f cycles/mma\n", ms * 1e-3 * 1.9e9 / iters); }
  return 0;
}
