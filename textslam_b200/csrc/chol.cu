// Dense FP64 Cholesky solve of the reduced camera system S y = b on one B200 (sm_100a).
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
// Per 64-wide panel j: potrf (one CTA, shared memory) -> trsm (one CTA per 64-row tile below) ->
// syrk/gemm trailing update with FP64 tensor-core MMA (mma.sync.m8n8k4.f64; tcgen05 has no FP64
// kind, DMMA is the FP64 tensor path on sm_100a), one CTA per 64x64 lower tile.
#include "ctx.cuh"
#include "solver.cuh"

namespace tsl {

constexpr int NB = 64;
constexpr int SPAD = NB + 4;  // smem row stride (doubles): conflict-free m8n8k4 fragment loads

// ---------------------------------------------------------------------------------------------
// potrf: factor the 64x64 diagonal tile in shared memory (right-looking, column by column).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) potrf_tile_kernel(double* __restrict__ A, int ld, int j, int* __restrict__ fail) {
  __shared__ double s[NB][NB + 1];
  double* Ajj = A + (size_t)j * NB * ld + (size_t)j * NB;
  for (int e = threadIdx.x; e < NB * NB; e += 256) {
    const int r = e >> 6, c = e & 63;
    s[r][c] = (c <= r) ? Ajj[(size_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  for (int c = 0; c < NB; ++c) {
    const double piv = s[c][c];
    if (!(piv > 0.0)) {  // not positive definite (or NaN): report and keep going with a harmless pivot
      if (threadIdx.x == 0) atomicExch(fail, 1);
    }
    const double d = (piv > 0.0) ? sqrt(piv) : 1.0;
    __syncthreads();
    const double inv = 1.0 / d;
    for (int r = c + threadIdx.x; r < NB; r += 256) s[r][c] = (r == c) ? d : s[r][c] * inv;
    __syncthreads();
    // trailing update: s[r][k] -= s[r][c]*s[k][c] for c < k <= r
    const int m = NB - 1 - c;  // trailing dimension
    for (int e = threadIdx.x; e < m * m; e += 256) {
      const int rr = e / m, kk = e - rr * m;
      if (kk <= rr) {
        const int r = c + 1 + rr, k = c + 1 + kk;
        s[r][k] -= s[r][c] * s[k][c];
      }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < NB * NB; e += 256) {
    const int r = e >> 6, c = e & 63;
    if (c <= r) Ajj[(size_t)r * ld + c] = s[r][c];
  }
}

// ---------------------------------------------------------------------------------------------
// trsm: X L_jj^T = A_ij for every 64-row tile i > j (one thread per row, x kept in registers).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB) trsm_tile_kernel(double* __restrict__ A, int ld, int j, int rows_valid_last) {
  __shared__ double sL[NB][NB + 1];
  const double* Ljj = A + (size_t)j * NB * ld + (size_t)j * NB;
  for (int e = threadIdx.x; e < NB * NB; e += NB) {
    const int r = e >> 6, c = e & 63;
    sL[r][c] = (c <= r) ? Ljj[(size_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  const int i = j + 1 + blockIdx.x;
  double* row = A + ((size_t)i * NB + threadIdx.x) * ld + (size_t)j * NB;
  double x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = row[c];
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double s = x[c];
#pragma unroll
    for (int k = 0; k < c; ++k) s -= x[k] * sL[c][k];
    x[c] = s / sL[c][c];
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) row[c] = x[c];
}

// ---------------------------------------------------------------------------------------------
// syrk / gemm trailing update with DMMA: A_ik -= X_i X_k^T for j < k <= i.
// CTA = 4 warps (2x2), warp tile 32x32 = 4x4 m8n8k4 tiles.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) syrk_tile_kernel(double* __restrict__ A, int ld, int j, int nt /*tiles below j*/, int skip_last_diag) {
  extern __shared__ double smem[];
  double* sA = smem;               // X_i  [64][SPAD]
  double* sB = smem + NB * SPAD;   // X_k  [64][SPAD]
  // decode lower-triangular tile index -> (ti >= tk)
  int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while (ti * (ti + 1) / 2 > t) --ti;
  const int tk = t - ti * (ti + 1) / 2;
  if (skip_last_diag && ti == nt - 1 && tk == nt - 1) return;  // (b-row, b-row) tile is never read
  const int gi = j + 1 + ti, gk = j + 1 + tk;
  const double* Xi = A + (size_t)gi * NB * ld + (size_t)j * NB;
  const double* Xk = A + (size_t)gk * NB * ld + (size_t)j * NB;
  for (int e = threadIdx.x; e < NB * NB / 2; e += 128) {  // double2 loads
    const int r = e >> 5, c2 = (e & 31) * 2;
    const double2 va = *reinterpret_cast<const double2*>(Xi + (size_t)r * ld + c2);
    const double2 vb = *reinterpret_cast<const double2*>(Xk + (size_t)r * ld + c2);
    sA[r * SPAD + c2] = va.x; sA[r * SPAD + c2 + 1] = va.y;
    sB[r * SPAD + c2] = vb.x; sB[r * SPAD + c2 + 1] = vb.y;
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
  double* C = A + (size_t)gi * NB * ld + (size_t)gk * NB;
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
// backward solve L^T x = y, right-looking over panels j = Tn-1 .. 0. y lives in `x` (in/out).
// CTA k < j: t = L_jk^T x_j ; y_k -= t.  Every CTA first solves L_jj^T x_j = y_j redundantly;
// CTA 0 (or the only CTA when j == 0) publishes x_j.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB) backsolve_panel_kernel(const double* __restrict__ A, int ld, int j, double* __restrict__ x,
                                                             double* __restrict__ xout) {
  __shared__ double sL[NB][NB + 1];
  __shared__ double sx[NB];
  const double* Ljj = A + (size_t)j * NB * ld + (size_t)j * NB;
  for (int e = threadIdx.x; e < NB * NB; e += NB) {
    const int r = e >> 6, c = e & 63;
    sL[r][c] = (c <= r) ? Ljj[(size_t)r * ld + c] : 0.0;
  }
  sx[threadIdx.x] = x[j * NB + threadIdx.x];
  __syncthreads();
  // column-oriented back substitution: x_c = y_c / L_cc ; y_r -= L_cr x_c (r < c)
  for (int c = NB - 1; c >= 0; --c) {
    if (threadIdx.x == c) sx[c] = sx[c] / sL[c][c];
    __syncthreads();
    if (threadIdx.x < c) sx[threadIdx.x] -= sL[c][threadIdx.x] * sx[c];
    __syncthreads();
  }
  const int k = (int)blockIdx.x - 1;  // block 0 publishes x_j, blocks 1..j update y_{k}
  if (k < 0) { xout[j * NB + threadIdx.x] = sx[threadIdx.x]; return; }
  // t[col] = sum_r L_jk[r][col] * x_j[r]; thread = col -> coalesced row reads
  const double* Ljk = A + (size_t)j * NB * ld + (size_t)k * NB;
  double t = 0.0;
#pragma unroll 8
  for (int r = 0; r < NB; ++r) t += Ljk[(size_t)r * ld + threadIdx.x] * sx[r];
  x[k * NB + threadIdx.x] -= t;
}

// gather y from the b row, scatter of x back handled by caller
__global__ void copy_row_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

int chol_workspace_dims(int n, int* ld, int* rows) {
  const int Tn = (n + NB - 1) / NB;
  *ld = Tn * NB;
  *rows = (Tn + 1) * NB;
  return Tn;
}

// Factor + solve. A: (Tn+1)*64 x ld as described above. ywork: ld doubles scratch. xout: ld doubles.
int chol_solve(tslam_ctx* ctx, double* A, int n, double* ywork, double* xout, int* d_fail) {
  int ld, rows;
  const int Tn = chol_workspace_dims(n, &ld, &rows);
  static bool attr_set = false;
  const int smem = 2 * NB * SPAD * (int)sizeof(double);
  if (!attr_set) {
    TSL_CUDA(cudaFuncSetAttribute(syrk_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  cudaStream_t s = ctx->stream;
  for (int j = 0; j < Tn; ++j) {
    potrf_tile_kernel<<<1, 256, 0, s>>>(A, ld, j, d_fail);
    const int nt = Tn - j;  // row tiles below the diagonal one, including the b tile row
    trsm_tile_kernel<<<nt, NB, 0, s>>>(A, ld, j, 0);
    if (nt > 0) syrk_tile_kernel<<<nt * (nt + 1) / 2, 128, smem, s>>>(A, ld, j, nt, 1);
  }
  TSL_CHECK_LAUNCH();
  copy_row_kernel<<<(ld + 255) / 256, 256, 0, s>>>(A + (size_t)Tn * NB * ld, ywork, ld);
  for (int j = Tn - 1; j >= 0; --j) backsolve_panel_kernel<<<j + 1, NB, 0, s>>>(A, ld, j, ywork, xout);
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl
