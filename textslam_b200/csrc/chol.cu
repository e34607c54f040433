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
// Tn x Tn tile pattern once per problem (ChoSymbolic) and every panel step only touches the listed
// non-zero tiles; a dense pattern degenerates to the classic right-looking blocked algorithm.
//
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
// device tile routines (potrf / trsm / invert: 64 threads; gemm: 128 threads)
// ---------------------------------------------------------------------------------------------

// The three 64x64 tile routines keep the tile in shared memory and are written column-oriented
// (right-looking) so that every step is a batch of INDEPENDENT FMAs: measured alternatives were a fully
// unrolled register version (64 us per launch: 3 x 2016 FMAs of straight-line code miss the instruction
// cache) and a left-looking dot-product version (101 us: serial LDS->DFMA chains) — profiles/r1_notes.md.
constexpr int LDT = NB + 1;       // smem leading dimension (doubles)
constexpr int PT_THREADS = 256;   // CTA size of potrf_trsm_kernel

// Right-looking Cholesky in place on sT (lower triangle), 256 threads as a 16x16 grid, each owning the
// elements (r, k) with r = ty + 16 i, k = tx + 16 j. sinv[c] = 1 / L[c][c]; col[] is a scaled copy of the
// current column so that the rank-1 update needs no read-after-write hazard handling.
__device__ __forceinline__ void potrf_tile(double* sT, double* sinv, double* col, int* fail) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int c = 0; c < NB; ++c) {
    const double piv = sT[c * LDT + c];                   // broadcast read; every thread derives the same inverse
    const double inv = (piv > 0.0) ? rsqrt(piv) : 1.0;
    if (threadIdx.x < NB) {
      const int r = threadIdx.x;
      if (r == c) { sinv[c] = inv; if (!(piv > 0.0)) atomicExch(fail, 1); }   // not SPD / NaN: report, harmless pivot
      if (r >= c) col[r] = sT[r * LDT + c] * inv;          // sT itself is written after the barrier (others still read sT[c][c])
    }
    __syncthreads();
    if (threadIdx.x < NB && threadIdx.x >= c) sT[threadIdx.x * LDT + c] = col[threadIdx.x];
    // trailing update A[r][k] -= l_r l_k for c < k <= r. All loads are staged in registers before any store so
    // that the 16 updates are independent (sT / col may alias as far as the compiler knows).
    const int k0 = c + 1;
    double lr[4], lk[4], a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { lr[i] = col[ty + 16 * i]; lk[i] = col[tx + 16 * i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ty + 16 * i, k = tx + 16 * j;
        a[i][j] = (k >= k0 && k <= r) ? sT[r * LDT + k] : 0.0;
      }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = ty + 16 * i, k = tx + 16 * j;
        if (k >= k0 && k <= r) sT[r * LDT + k] = a[i][j] - lr[i] * lk[j];
      }
    __syncthreads();
  }
}

// X L^T = A for one 64-row tile held in sX (in place), column-oriented: x_c *= inv_c, then x_k -= x_c L[k][c]
// for k > c. Four lanes of one warp share a row (k = q + 4 m); rows are independent, so only __syncwarp is needed.
__device__ __forceinline__ void trsm_tile(double* sX, const double* sT, const double* sinv) {
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
  double* x = sX + r * LDT;
  for (int c = 0; c < NB; ++c) {
    const double xc = x[c] * sinv[c];
    __syncwarp();
    if (q == (c & 3)) x[c] = xc;
    const int kb = c + 1 + ((q - (c + 1)) & 3);
    double xv[16], lv[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) { const int k = kb + 4 * m; xv[m] = (k < NB) ? x[k] : 0.0; lv[m] = (k < NB) ? sT[k * LDT + c] : 0.0; }
#pragma unroll
    for (int m = 0; m < 16; ++m) { const int k = kb + 4 * m; if (k < NB) x[k] = xv[m] - xc * lv[m]; }
    __syncwarp();
  }
}

// L^-1 for the backward solve comes from the same routine: CTA 0 runs trsm_tile on an identity tile
// (X L^T = I  ->  X = L^-T) and stores the transpose.

__device__ __forceinline__ void load_L_tile(const double* __restrict__ Ljj, int ld, double (*sL)[NB + 1]) {
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e >> 6, c = e & 63;
    sL[r][c] = (c <= r) ? Ljj[(size_t)r * ld + c] : 0.0;
  }
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
__global__ void __launch_bounds__(PT_THREADS) potrf_trsm_kernel(double* __restrict__ A, int ld, int j, const int* __restrict__ rows,
                                                                int* __restrict__ fail, double* __restrict__ Linv) {
  extern __shared__ double smem[];
  double* sT = smem;                 // 64 x LDT: diagonal tile -> L_jj
  double* sX = smem + NB * LDT;      // 64 x LDT: row tile (TRSM) or identity -> L_jj^-T (CTA 0)
  __shared__ double sinv[NB];
  __shared__ double col[NB];
  const double* Ajj = A + (size_t)j * NB * ld + (size_t)j * NB;
  double* Aij = blockIdx.x == 0 ? nullptr : A + (size_t)rows[blockIdx.x - 1] * NB * ld + (size_t)j * NB;
  {
    double vt[16], vx[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = threadIdx.x + PT_THREADS * u, r = e >> 6, c = e & 63;
      vt[u] = (c <= r) ? Ajj[(size_t)r * ld + c] : 0.0;
      vx[u] = Aij ? Aij[(size_t)r * ld + c] : ((r == c) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int e = threadIdx.x + PT_THREADS * u, r = e >> 6, c = e & 63;
      sT[r * LDT + c] = vt[u]; sX[r * LDT + c] = vx[u];
    }
  }
  __syncthreads();
  potrf_tile(sT, sinv, col, fail);
  trsm_tile(sX, sT, sinv);           // CTA 0: X L^T = I  ->  X = L^-T
  __syncthreads();
  if (blockIdx.x == 0) {             // off the critical path: store L_jj^-1 = X^T for the backward solve
    double* dst = Linv + (size_t)j * NB * NB;
    for (int e = threadIdx.x; e < NB * NB; e += PT_THREADS) { const int r = e >> 6, c = e & 63; dst[e] = sX[c * LDT + r]; }
    return;
  }
  for (int e = threadIdx.x; e < NB * NB; e += PT_THREADS) { const int r = e >> 6, c = e & 63; Aij[(size_t)r * ld + c] = sX[r * LDT + c]; }
}

// trailing update: for each listed pair (i,k), i >= k > j: A_ik -= X_i X_k^T
__global__ void __launch_bounds__(128) syrk_pairs_kernel(double* __restrict__ A, int ld, int j, const int2* __restrict__ pairs) {
  extern __shared__ double smem[];
  const int2 pr = pairs[blockIdx.x];
  gemm_tile_nt(A + (size_t)pr.x * NB * ld + (size_t)j * NB, A + (size_t)pr.y * NB * ld + (size_t)j * NB,
               A + (size_t)pr.x * NB * ld + (size_t)pr.y * NB, ld, smem, smem + NB * SPAD);
}

// backward solve L^T x = y, right-looking over panels j = Tn-1 .. 0. y lives in `x` (in/out).
// x_j = L_jj^-T y_j is a 64x64 mat-vec with the stored (L_jj^-1)^T (no substitution chain); CTA 0
// publishes x_j, CTA 1+m applies y_k -= L_jk^T x_j for the m-th non-zero tile (j,k), k < j.
__global__ void __launch_bounds__(NB) backsolve_panel_kernel(const double* __restrict__ A, int ld, int j, const int* __restrict__ cols,
                                                             const double* __restrict__ LinvT, double* __restrict__ x,
                                                             double* __restrict__ xout) {
  __shared__ double sy[NB];
  __shared__ double sx[NB];
  sy[threadIdx.x] = x[j * NB + threadIdx.x];
  __syncthreads();
  {
    // x_c = sum_r (L^-1)[r][c] y_r : thread c walks column c, rows are contiguous across threads (coalesced)
    const double* M = LinvT + (size_t)j * NB * NB;
    double t0 = 0.0, t1 = 0.0;
#pragma unroll 8
    for (int r = 0; r < NB; r += 2) {
      t0 += M[r * NB + threadIdx.x] * sy[r];
      t1 += M[(r + 1) * NB + threadIdx.x] * sy[r + 1];
    }
    sx[threadIdx.x] = t0 + t1;
  }
  __syncthreads();
  if (blockIdx.x == 0) { xout[j * NB + threadIdx.x] = sx[threadIdx.x]; return; }
  const int k = cols[blockIdx.x - 1];
  // t[col] = sum_r L_jk[r][col] * x_j[r]; thread = col -> coalesced row reads
  const double* Ljk = A + (size_t)j * NB * ld + (size_t)k * NB;
  double t0 = 0.0, t1 = 0.0;
#pragma unroll 8
  for (int r = 0; r < NB; r += 2) {
    t0 += Ljk[(size_t)r * ld + threadIdx.x] * sx[r];
    t1 += Ljk[(size_t)(r + 1) * ld + threadIdx.x] * sx[r + 1];
  }
  x[k * NB + threadIdx.x] -= t0 + t1;
}

__global__ void copy_row_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// host: symbolic tile factorisation + launch sequence
// ---------------------------------------------------------------------------------------------
int chol_workspace_dims(int n, int* ld, int* rows) {
  const int Tn = (n + NB - 1) / NB;
  *ld = Tn * NB;
  *rows = (Tn + 1) * NB;
  return Tn;
}

// tile_nz: Tn x Tn row-major flags of the lower-triangular tile pattern of S (diagonal always set).
int chol_symbolic(tslam_ctx* ctx, int n, const std::vector<uint8_t>& tile_nz, CholSymbolic* sym) {
  int ld, rows;
  const int Tn = chol_workspace_dims(n, &ld, &rows);
  sym->Tn = Tn; sym->n = n;
  const int T1 = Tn + 1;  // + the b tile row (dense)
  std::vector<uint8_t> P((size_t)T1 * T1, 0);
  for (int i = 0; i < Tn; ++i)
    for (int k = 0; k <= i; ++k) P[(size_t)i * T1 + k] = (i == k) || tile_nz[(size_t)i * Tn + k];
  for (int k = 0; k < Tn; ++k) P[(size_t)Tn * T1 + k] = 1;
  std::vector<int> rows_h, rows_ptr(Tn + 1, 0), cols_h, cols_ptr(Tn + 1, 0), pairs_ptr(Tn + 1, 0);
  std::vector<int2> pairs_h;
  std::vector<int> nzrows;
  long long flop_tiles = 0;
  for (int j = 0; j < Tn; ++j) {
    nzrows.clear();
    for (int i = j + 1; i < T1; ++i) if (P[(size_t)i * T1 + j]) nzrows.push_back(i);
    for (int i : nzrows) rows_h.push_back(i);
    rows_ptr[j + 1] = (int)rows_h.size();
    for (size_t a = 0; a < nzrows.size(); ++a)
      for (size_t b = 0; b <= a; ++b) {
        const int i = nzrows[a], k = nzrows[b];
        if (i == Tn && k == Tn) continue;  // (b row, b row) is never read
        P[(size_t)i * T1 + k] = 1;         // fill
        pairs_h.push_back(make_int2(i, k));
      }
    pairs_ptr[j + 1] = (int)pairs_h.size();
    flop_tiles += (long long)(pairs_ptr[j + 1] - pairs_ptr[j]);
  }
  // backward solve: for panel j the non-zero tiles (j,k), k < j of the FACTOR
  for (int j = 0; j < Tn; ++j) {
    for (int k = 0; k < j; ++k) if (P[(size_t)j * T1 + k]) cols_h.push_back(k);
    cols_ptr[j + 1] = (int)cols_h.size();
  }
  sym->rows_ptr = rows_ptr; sym->pairs_ptr = pairs_ptr; sym->cols_ptr = cols_ptr;
  sym->gemm_tiles = flop_tiles;
  cudaStream_t s = ctx->stream;
  TSL_CUDA(sym->rows.upload(rows_h.data(), rows_h.size(), s));
  TSL_CUDA(sym->pairs.upload(pairs_h.data(), pairs_h.size(), s));
  TSL_CUDA(sym->cols.upload(cols_h.data(), cols_h.size(), s));
  TSL_CUDA(sym->Ldiag.reserve((size_t)(Tn ? Tn : 1) * NB * NB));
  TSL_CUDA(cudaStreamSynchronize(s));
  return TSLAM_OK;
}

// Factor + solve. A: (Tn+1)*64 x ld as described above. ywork: ld doubles scratch. xout: ld doubles.
int chol_solve(tslam_ctx* ctx, const CholSymbolic& sym, double* A, double* ywork, double* xout, int* d_fail) {
  int ld, rows;
  const int Tn = chol_workspace_dims(sym.n, &ld, &rows);
  static bool attr_set = false;
  const int smem = 2 * NB * SPAD * (int)sizeof(double);
  const int smem_pt = 2 * NB * LDT * (int)sizeof(double);
  if (!attr_set) {
    TSL_CUDA(cudaFuncSetAttribute(syrk_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TSL_CUDA(cudaFuncSetAttribute(potrf_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pt));
    attr_set = true;
  }
  cudaStream_t s = ctx->stream;
  for (int j = 0; j < Tn; ++j) {
    const int nrows = sym.rows_ptr[j + 1] - sym.rows_ptr[j];
    const int npairs = sym.pairs_ptr[j + 1] - sym.pairs_ptr[j];
    LAUNCH(potrf_trsm_kernel<<<1 + nrows, PT_THREADS, smem_pt, s>>>(A, ld, j, sym.rows.p + sym.rows_ptr[j], d_fail, sym.Ldiag.p));
    if (npairs > 0) LAUNCH(syrk_pairs_kernel<<<npairs, 128, smem, s>>>(A, ld, j, sym.pairs.p + sym.pairs_ptr[j]));
  }
  TSL_CHECK_LAUNCH();
  LAUNCH(copy_row_kernel<<<(ld + 255) / 256, 256, 0, s>>>(A + (size_t)Tn * NB * ld, ywork, ld));
  for (int j = Tn - 1; j >= 0; --j) {
    const int ncols = sym.cols_ptr[j + 1] - sym.cols_ptr[j];
    LAUNCH(backsolve_panel_kernel<<<1 + ncols, NB, 0, s>>>(A, ld, j, sym.cols.p + sym.cols_ptr[j], sym.Ldiag.p, ywork, xout));
  }
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl
