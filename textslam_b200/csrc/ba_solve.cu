// Levenberg-Marquardt solver with landmark marginalisation (Schur complement) on one or more B200s.
//
// Replaces `ceres::Solve` + `Problem::Evaluate` at src/optimizer.cc:1222-1233 (pose-only),
// :1602-1614 (local BA), :1840 (global BA), :1982 (landmarks), :2209 (theta). The trust-region loop
// follows Ceres' published TrustRegionMinimizer / LevenbergMarquardtStrategy behaviour
// (Ceres is not under /root/reference) and is mirrored by oracle/ba_lm.cpp.
//
// Per LM iteration on the device:
//   eval (ba_eval.cu)   r, J per observation, observation-major, Huber-corrected
//   accum               per landmark  V = J_l'J_l, g_l ; per (landmark,camera) slot E = J_c'J_l
//   vinv                (V + D^2)^-1 for the current radius
//   block               one warp per non-zero 6x6 block of the reduced camera matrix: gathers the
//                       J_a'J_b terms and the -E_a V^-1 E_b' terms from precomputed entry lists
//                       (deterministic, no atomics), warp-shuffle reduction, plus b and the raw gradient
//   [all-reduce]        one NCCL sum over [blocks | b | gradient | costs] (multi-GPU global BA)
//   scatter + Cholesky  dense FP64 blocked Cholesky (chol.cu), forward solve folded in
//   backsub             landmark steps, candidate parameters (Ceres Plus), step norms
//   model / candidate   model cost change -(Jd)'(r + Jd/2) and the candidate cost
// The host only reads back a handful of scalars per iteration to take the accept / reject decision.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>
#include <atomic>
#include "ctx.cuh"
#include "solver.cuh"
#include "analysis.hpp"
#include "ba_device.cuh"

namespace tsl {

// ------------------------------------------------------------------------------------------------
// scalar slots (device, doubles). [0..7] are summed across ranks, mx[] is max-reduced.
// ------------------------------------------------------------------------------------------------
enum { SC_COST = 0, SC_FIXED = 1, SC_CAND = 2, SC_CAND_FIXED = 3, SC_MCC = 4, SC_STEP2 = 5, SC_CNORM2 = 6, SC_XNORM2 = 7, SC_N = 8 };
enum { MX_GMAX = 0, MX_FAIL = 1, MX_N = 2 };

// ================================================================================================
// kernels
// ================================================================================================

// deterministic two-level sum: parts[0..n) (stride 1) -> *out (+= if accumulate)
__global__ void __launch_bounds__(256) sum_parts_kernel(const double* __restrict__ parts, int n, int stride, int offset, double* out, int accumulate) {
  PDL_PROLOGUE();
  __shared__ double s[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += parts[(size_t)i * stride + offset];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = accumulate ? *out + s[0] : s[0];
}

// two interleaved series (stride 2) -> two scalars in one launch: CTA b sums parts[2 i + b]; same order as sum_parts_kernel
__global__ void __launch_bounds__(256) sum_parts_pair_kernel(const double* __restrict__ parts, int n, double* out0, double* out1, int accumulate) {
  PDL_PROLOGUE();
  __shared__ double s[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += parts[(size_t)i * 2 + blockIdx.x];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  double* out = blockIdx.x ? out1 : out0;
  if (threadIdx.x == 0) *out = accumulate ? *out + s[0] : s[0];
}

__device__ __forceinline__ double block_sum_256(double v, double* s) {
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  return s[0];
}

// ---- per-landmark accumulation: V (DxD), g (D), both in the Jacobi-scaled system ------------------
template <int D, int ROWS, int JC>
__device__ __forceinline__ void lm_accum_body(const unsigned bid, int nv, const int* __restrict__ obs_ptr, const int* __restrict__ obs, const double* __restrict__ J,
                                const double* __restrict__ r, const double* __restrict__ scale, double* __restrict__ V,
                                double* __restrict__ g) {
  const int v = bid * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  double Vv[D * D], gv[D];
#pragma unroll
  for (int k = 0; k < D * D; ++k) Vv[k] = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) gv[k] = 0.0;
  for (int e = obs_ptr[v]; e < obs_ptr[v + 1]; ++e) {
    const int i = obs[e];
    const double* Ji = J + (size_t)i * ROWS * JC;
    const double* ri = r + (size_t)i * ROWS;
#pragma unroll
    for (int row = 0; row < ROWS; ++row) {
      double jl[D];
#pragma unroll
      for (int a = 0; a < D; ++a) jl[a] = Ji[row * JC + 12 + a];
      const double rr = ri[row];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        gv[a] += jl[a] * rr;
#pragma unroll
        for (int b = 0; b < D; ++b) Vv[a * D + b] += jl[a] * jl[b];
      }
    }
  }
  double s[D];
#pragma unroll
  for (int a = 0; a < D; ++a) s[a] = scale[v * D + a];
#pragma unroll
  for (int a = 0; a < D; ++a) {
    g[v * D + a] = gv[a] * s[a];
#pragma unroll
    for (int b = 0; b < D; ++b) V[(size_t)v * D * D + a * D + b] = Vv[a * D + b] * s[a] * s[b];
  }
}
template <int D, int ROWS, int JC>
__global__ void lm_accum_kernel(int nv, const int* __restrict__ obs_ptr, const int* __restrict__ obs, const double* __restrict__ J,
                                const double* __restrict__ r, const double* __restrict__ scale, double* __restrict__ V,
                                double* __restrict__ g) {
  PDL_PROLOGUE();
  lm_accum_body<D, ROWS, JC>(blockIdx.x, nv, obs_ptr, obs, J, r, scale, V, g);
}

// ---- per (landmark, camera) slot: E = S_c J_c' J_l S_l  (6 x D) ----------------------------------------
template <int D, int ROWS, int JC>
__device__ __forceinline__ void slot_accum_body(const unsigned bid, int ns, const int* __restrict__ ent_ptr, const int* __restrict__ ent, const int* __restrict__ slot_cam,
                                  const int* __restrict__ slot_lm, const double* __restrict__ J, const double* __restrict__ scale_c,
                                  const double* __restrict__ scale_l, double* __restrict__ E) {
  const int sidx = bid * blockDim.x + threadIdx.x;
  if (sidx >= ns) return;
  double Ev[6 * D];
#pragma unroll
  for (int k = 0; k < 6 * D; ++k) Ev[k] = 0.0;
  for (int e = ent_ptr[sidx]; e < ent_ptr[sidx + 1]; ++e) {
    const int code = ent[e];
    const int i = code >> 1, off = (code & 1) * 6;
    const double* Ji = J + (size_t)i * ROWS * JC;
#pragma unroll
    for (int row = 0; row < ROWS; ++row) {
      double jl[D];
#pragma unroll
      for (int a = 0; a < D; ++a) jl[a] = Ji[row * JC + 12 + a];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double jc = Ji[row * JC + off + c];
#pragma unroll
        for (int a = 0; a < D; ++a) Ev[c * D + a] += jc * jl[a];
      }
    }
  }
  const int cam = slot_cam[sidx], lm = slot_lm[sidx];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double sc = scale_c[6 * cam + c];
#pragma unroll
    for (int a = 0; a < D; ++a) E[(size_t)sidx * 6 * D + c * D + a] = Ev[c * D + a] * sc * scale_l[lm * D + a];
  }
}
template <int D, int ROWS, int JC>
__global__ void slot_accum_kernel(int ns, const int* __restrict__ ent_ptr, const int* __restrict__ ent, const int* __restrict__ slot_cam,
                                  const int* __restrict__ slot_lm, const double* __restrict__ J, const double* __restrict__ scale_c,
                                  const double* __restrict__ scale_l, double* __restrict__ E) {
  PDL_PROLOGUE();
  slot_accum_body<D, ROWS, JC>(blockIdx.x, ns, ent_ptr, ent, slot_cam, slot_lm, J, scale_c, scale_l, E);
}

// Same sums with one WARP per landmark (lanes over the (observation, residual row) pairs, butterfly at the end): the text
// planes are few (tens) with hundreds of residual rows each — a thread per plane left the GPU idle behind 30 serial loops
// (local BA C4: 94 us per call, profiles/r1_notes.md).
template <int D, int ROWS, int JC>
__device__ __forceinline__ void lm_accum_warp_body(const unsigned bid, int nv, const int* __restrict__ obs_ptr, const int* __restrict__ obs, const double* __restrict__ J,
                                                            const double* __restrict__ r, const double* __restrict__ scale, double* __restrict__ V,
                                                            double* __restrict__ g) {
  const int v = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (v >= nv) return;   // warp-uniform
  double Vv[D * D], gv[D];
#pragma unroll
  for (int k = 0; k < D * D; ++k) Vv[k] = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) gv[k] = 0.0;
  const int e0 = obs_ptr[v], n_items = (obs_ptr[v + 1] - e0) * ROWS;
  for (int t = lane; t < n_items; t += 32) {
    const int i = obs[e0 + t / ROWS], row = t % ROWS;
    const double* Ji = J + (size_t)i * ROWS * JC + row * JC + 12;
    const double rr = r[(size_t)i * ROWS + row];
    double jl[D];
#pragma unroll
    for (int a = 0; a < D; ++a) jl[a] = Ji[a];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      gv[a] += jl[a] * rr;
#pragma unroll
      for (int b = 0; b < D; ++b) Vv[a * D + b] += jl[a] * jl[b];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < D * D; ++k) Vv[k] += __shfl_xor_sync(0xffffffffu, Vv[k], o);
#pragma unroll
    for (int k = 0; k < D; ++k) gv[k] += __shfl_xor_sync(0xffffffffu, gv[k], o);
  }
  if (lane == 0) {
    double s[D];
#pragma unroll
    for (int a = 0; a < D; ++a) s[a] = scale[v * D + a];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      g[v * D + a] = gv[a] * s[a];
#pragma unroll
      for (int b = 0; b < D; ++b) V[(size_t)v * D * D + a * D + b] = Vv[a * D + b] * s[a] * s[b];
    }
  }
}
template <int D, int ROWS, int JC>
__global__ void __launch_bounds__(128) lm_accum_warp_kernel(int nv, const int* __restrict__ obs_ptr, const int* __restrict__ obs, const double* __restrict__ J,
                                                            const double* __restrict__ r, const double* __restrict__ scale, double* __restrict__ V,
                                                            double* __restrict__ g) {
  PDL_PROLOGUE();
  lm_accum_warp_body<D, ROWS, JC>(blockIdx.x, nv, obs_ptr, obs, J, r, scale, V, g);
}

template <int D, int ROWS, int JC>
__device__ __forceinline__ void slot_accum_warp_body(const unsigned bid, int ns, const int* __restrict__ ent_ptr, const int* __restrict__ ent, const int* __restrict__ slot_cam,
                                                              const int* __restrict__ slot_lm, const double* __restrict__ J, const double* __restrict__ scale_c,
                                                              const double* __restrict__ scale_l, double* __restrict__ E) {
  const int sidx = (bid * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (sidx >= ns) return;   // warp-uniform
  double Ev[6 * D];
#pragma unroll
  for (int k = 0; k < 6 * D; ++k) Ev[k] = 0.0;
  const int e0 = ent_ptr[sidx], n_items = (ent_ptr[sidx + 1] - e0) * ROWS;
  for (int t = lane; t < n_items; t += 32) {
    const int code = ent[e0 + t / ROWS], row = t % ROWS;
    const int i = code >> 1, off = (code & 1) * 6;
    const double* Ji = J + (size_t)i * ROWS * JC + row * JC;
    double jl[D];
#pragma unroll
    for (int a = 0; a < D; ++a) jl[a] = Ji[12 + a];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double jc = Ji[off + c];
#pragma unroll
      for (int a = 0; a < D; ++a) Ev[c * D + a] += jc * jl[a];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 6 * D; ++k) Ev[k] += __shfl_xor_sync(0xffffffffu, Ev[k], o);
  const int cam = slot_cam[sidx], lm = slot_lm[sidx];
#pragma unroll
  for (int k = 0; k < 6 * D; ++k)
    if (lane == k) E[(size_t)sidx * 6 * D + k] = Ev[k] * scale_c[6 * cam + k / D] * scale_l[lm * D + k % D];
}
template <int D, int ROWS, int JC>
__global__ void __launch_bounds__(128) slot_accum_warp_kernel(int ns, const int* __restrict__ ent_ptr, const int* __restrict__ ent, const int* __restrict__ slot_cam,
                                                              const int* __restrict__ slot_lm, const double* __restrict__ J, const double* __restrict__ scale_c,
                                                              const double* __restrict__ scale_l, double* __restrict__ E) {
  PDL_PROLOGUE();
  slot_accum_warp_body<D, ROWS, JC>(blockIdx.x, ns, ent_ptr, ent, slot_cam, slot_lm, J, scale_c, scale_l, E);
}

// V, g per landmark and E per (landmark, camera) slot are independent sums over the same Jacobian: one launch, the first
// g_lm CTAs run the landmark bodies, the rest the slot bodies (two dependent-in-stream launches would run back to back).
template <int D, int ROWS, int JC, bool WARP>
__global__ void __launch_bounds__(128) accum_merged_kernel(int g_lm, int nv, const int* __restrict__ obs_ptr, const int* __restrict__ obs,
                                                           const double* __restrict__ J, const double* __restrict__ r, const double* __restrict__ scale_l,
                                                           double* __restrict__ V, double* __restrict__ g, int ns, const int* __restrict__ ent_ptr,
                                                           const int* __restrict__ ent, const int* __restrict__ slot_cam, const int* __restrict__ slot_lm,
                                                           const double* __restrict__ scale_c, double* __restrict__ E) {
  PDL_PROLOGUE();
  if ((int)blockIdx.x < g_lm) {
    if (WARP) lm_accum_warp_body<D, ROWS, JC>(blockIdx.x, nv, obs_ptr, obs, J, r, scale_l, V, g);
    else lm_accum_body<D, ROWS, JC>(blockIdx.x, nv, obs_ptr, obs, J, r, scale_l, V, g);
  } else {
    if (WARP) slot_accum_warp_body<D, ROWS, JC>(blockIdx.x - g_lm, ns, ent_ptr, ent, slot_cam, slot_lm, J, scale_c, scale_l, E);
    else slot_accum_body<D, ROWS, JC>(blockIdx.x - g_lm, ns, ent_ptr, ent, slot_cam, slot_lm, J, scale_c, scale_l, E);
  }
}

__device__ __forceinline__ double lm_damp(double d, double inv_radius) { return fmin(fmax(d, 1e-6), 1e32) * inv_radius; }

// ---- (V + D^2)^-1 for the current trust-region radius ----------------------------------------------
template <int D>
__global__ void lm_vinv_kernel(int nv, const double* __restrict__ V, double inv_radius, double* __restrict__ Vinv, double* __restrict__ mx) {
  PDL_PROLOGUE();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  if (D == 1) {
    const double d = V[v];
    Vinv[v] = 1.0 / (d + lm_damp(d, inv_radius));
  } else {
    const double* A = V + (size_t)v * 9;
    const double a = A[0] + lm_damp(A[0], inv_radius), b = A[1], c = A[2];
    const double d = A[4] + lm_damp(A[4], inv_radius), e = A[5], f = A[8] + lm_damp(A[8], inv_radius);
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (!(det > 0.0) || !(a > 0.0)) atomicMax(reinterpret_cast<unsigned long long*>(mx + MX_FAIL), __double_as_longlong(1.0));
    const double id = 1.0 / det;
    double* o = Vinv + (size_t)v * 9;
    o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
    o[3] = o[1]; o[4] = (a * f - c * c) * id; o[5] = (b * c - a * e) * id;
    o[6] = o[2]; o[7] = o[5]; o[8] = (a * d - b * b) * id;
  }
}

// ---- Jacobi scaling (iteration 0): unscaled squared column norms ------------------------------------
// one warp per free camera: sum over the diagonal block's direct entries
struct BlockLists {
  const int* dp_ptr; const int* dp;   // direct point entries  (obs << 2 | code)
  const int* dt_ptr; const int* dt;   // direct text entries
  const int* sp_ptr; const int2* sp;  // schur point entries (slot_i, slot_j)
  const int* st_ptr; const int2* st;  // schur text entries
};

__device__ __forceinline__ void code_offsets(int code, int& ox, int& oy) {
  // 0: (c,c)  1: (h,h)  2: rows = cam cols = host  3: rows = host cols = cam
  ox = (code == 1 || code == 3) ? 6 : 0;
  oy = (code == 1 || code == 2) ? 6 : 0;
}

__global__ void cam_colnorm_kernel(int nc, const int* __restrict__ diag_blk, BlockLists L, const double* __restrict__ pJ,
                                   const double* __restrict__ tJ, double* __restrict__ out /*6nc*/) {
  PDL_PROLOGUE();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nc) return;
  const int blk = diag_blk[warp];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int e = L.dp_ptr[blk] + lane; e < L.dp_ptr[blk + 1]; e += 32) {
    const int code = L.dp[e]; int ox, oy; code_offsets(code & 3, ox, oy);
    const double* Ji = pJ + (size_t)(code >> 2) * 26;
#pragma unroll
    for (int row = 0; row < 2; ++row)
#pragma unroll
      for (int c = 0; c < 6; ++c) { const double j = Ji[row * 13 + ox + c]; acc[c] += j * j; }
  }
  for (int e = L.dt_ptr[blk] + lane; e < L.dt_ptr[blk + 1]; e += 32) {
    const int code = L.dt[e]; int ox, oy; code_offsets(code & 3, ox, oy);
    const double* Ji = tJ + (size_t)(code >> 2) * 120;
#pragma unroll
    for (int row = 0; row < 8; ++row)
#pragma unroll
      for (int c = 0; c < 6; ++c) { const double j = Ji[row * 15 + ox + c]; acc[c] += j * j; }
  }
#pragma unroll
  for (int c = 0; c < 6; ++c)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
  if (lane < 6) out[6 * warp + lane] = acc[lane];
}

__global__ void scale_from_norm_kernel(int n, const double* __restrict__ d2, double* __restrict__ scale) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scale[i] = 1.0 / (1.0 + sqrt(d2[i]));
}
// landmark scales from the diagonal of the (unscaled, scale == 1) V
template <int D>
__global__ void lm_scale_kernel(int nv, const double* __restrict__ V, double* __restrict__ scale) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv * D) return;
  const int v = i / D, a = i - v * D;
  scale[i] = 1.0 / (1.0 + sqrt(V[(size_t)v * D * D + a * D + a]));
}
__global__ void fill_kernel(double* p, int n, double v) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- camera-block Gram matrices of the text blocks -------------------------------------------------------------------------
// A text block has 8 rows: as a direct entry of the reduced system it costs 96 scattered 8-byte loads and 288 FMAs per lane in
// schur_block_body, three times per block (camera, host and cross block) — 131 us of the text-on iteration against 55 us without
// text. Its products do not depend on the trust-region radius, so they are formed once per Jacobian: one warp per block stages the
// 8 x 15 Jacobian + residuals and writes [J_c'J_c (36) | J_h'J_h (36) | J_c'J_h (36) | J_c'r (6) | J_h'r (6)] = 120 doubles, which the
// gather then reads as one contiguous 36-double record per entry.
constexpr int TG_STRIDE = 120;
__global__ void __launch_bounds__(128) text_gram_kernel(int n, const double* __restrict__ tJ, const double* __restrict__ tr, double* __restrict__ G) {
  PDL_PROLOGUE();
  __shared__ double s[4][128];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + w;
  if (b >= n) return;
  double* sj = s[w];
#pragma unroll
  for (int j = 0; j < 4; ++j) { const int q = lane + 32 * j; sj[q] = q < 120 ? tJ[(size_t)b * 120 + q] : tr[(size_t)b * 8 + q - 120]; }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = lane + 32 * j;
    if (o >= TG_STRIDE) break;
    int ia, ib;   // column of J (0..11), or 15 = the residual
    if (o < 36) { ia = o / 6; ib = o % 6; }
    else if (o < 72) { ia = 6 + (o - 36) / 6; ib = 6 + (o - 36) % 6; }
    else if (o < 108) { ia = (o - 72) / 6; ib = 6 + (o - 72) % 6; }
    else { ia = o - 108; ib = 15; }
    double v = 0.0;
#pragma unroll
    for (int row = 0; row < 8; ++row) v += sj[15 * row + ia] * (ib < 15 ? sj[15 * row + ib] : sj[120 + row]);
    G[(size_t)b * TG_STRIDE + o] = v;
  }
}

// ---- reduced camera system: one warp per non-zero upper block (a <= b) -------------------------------
struct BlockArgs {
  int nblk; const int* blk_a; const int* blk_b; BlockLists L;
  const double* pJ; const double* pr; const double* tJ; const double* tr;
  const double* tG;   // per text block: Gram matrices of its camera / host columns (text_gram_kernel)
  const double* scale_c;
  const double* Ep; const double* Vinvp; const double* gp; const int* sp_lm;
  const double* Et; const double* Vinvt; const double* gt; const int* st_lm;
  double* Sblk;   // nblk x 36, block (a,b) row-major: rows = a's tangent, cols = b's tangent (UNDAMPED)
  double* bvec;   // 6 nc : reduced right-hand side (scaled)
  double* graw;   // 6 nc : unscaled gradient J'r of the camera blocks
  double* udiag;  // 6 nc : diagonal of the scaled J_c'J_c (the LM damping term is built from it after the reduce)
};

// G lanes per block (G = 32 for the diagonal blocks, which carry ~650 gather entries each on the global-BA shape;
// G = 8 for the off-diagonal ones with ~35): the fixed cost of a block is the butterfly reduction of its 36 partial
// sums (36 x log2(G) 64-bit shuffles), which dominated when every block had a whole warp.
template <int G>
__device__ __forceinline__ void schur_block_body(const unsigned bid, BlockArgs A, const int* __restrict__ list, int nlist) {
  static_assert(G == 8 || G == 32 || G == 128, "group size");
  __shared__ double xs[G == 128 ? 4 * 54 : 1];   // G == 128 (one CTA per block): cross-warp stage of the reduction
  const int gi = (bid * blockDim.x + threadIdx.x) / G, lane = threadIdx.x & (G - 1);
  const bool valid = gi < nlist;           // lanes of an empty group still take part in the full-warp shuffles
  const int blk = valid ? (list ? list[gi] : gi) : 0;   // list == NULL: every block, in order
  const int a = A.blk_a[blk], b = A.blk_b[blk];
  const bool diag = a == b;
  double sa[6], sb[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) { sa[c] = A.scale_c[6 * a + c]; sb[c] = A.scale_c[6 * b + c]; }
  double acc[36], gr[6];
#pragma unroll
  for (int k = 0; k < 36; ++k) acc[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) gr[k] = 0.0;
  // ---- direct J_a' J_b terms (scaled on load) ----
  for (int e = A.L.dp_ptr[blk] + lane, e1 = valid ? A.L.dp_ptr[blk + 1] : 0; e < e1; e += G) {
    const int code = A.L.dp[e]; int ox, oy; code_offsets(code & 3, ox, oy);
    const int i = code >> 2;
    const double* Ji = A.pJ + (size_t)i * 26;
#pragma unroll
    for (int row = 0; row < 2; ++row) {
      double jx[6], jy[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) { jx[c] = Ji[row * 13 + ox + c] * sa[c]; jy[c] = Ji[row * 13 + oy + c] * sb[c]; }
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[p * 6 + q] += jx[p] * jy[q];
      if (diag) {
        const double rr = A.pr[(size_t)i * 2 + row];
#pragma unroll
        for (int p = 0; p < 6; ++p) gr[p] += jx[p] * rr;
      }
    }
  }
  for (int e = A.L.dt_ptr[blk] + lane, e1 = valid ? A.L.dt_ptr[blk + 1] : 0; e < e1; e += G) {
    const int code = A.L.dt[e], cd = code & 3;
    const int i = code >> 2;
    // 0: (c,c)  1: (h,h)  2: rows = cam, cols = host  3: rows = host, cols = cam (the transpose of the stored J_c'J_h)
    const double* Gm = A.tG + (size_t)i * TG_STRIDE + (cd == 0 ? 0 : (cd == 1 ? 36 : 72));
    if (cd == 3) {
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[p * 6 + q] += Gm[q * 6 + p] * (sa[p] * sb[q]);
    } else {
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[p * 6 + q] += Gm[p * 6 + q] * (sa[p] * sb[q]);
    }
    if (diag) {
      const double* gv = A.tG + (size_t)i * TG_STRIDE + (cd == 0 ? 108 : 114);
#pragma unroll
      for (int p = 0; p < 6; ++p) gr[p] += gv[p] * sa[p];
    }
  }
  double dd[6], bred[6];   // diagonal of the direct part (LM damping is built from it), Schur part of the right-hand side
#pragma unroll
  for (int k = 0; k < 6; ++k) { dd[k] = acc[7 * k]; bred[k] = 0.0; }
  // ---- Schur terms: - E_i V^-1 E_j' (and - E_i V^-1 g for the right-hand side), subtracted in place ----
  for (int e = A.L.sp_ptr[blk] + lane, e1 = valid ? A.L.sp_ptr[blk + 1] : 0; e < e1; e += G) {
    const int2 sl = A.L.sp[e];
    const double* Ei = A.Ep + (size_t)sl.x * 6;
    const double* Ej = A.Ep + (size_t)sl.y * 6;
    const int lm = A.sp_lm[sl.x];
    const double vi = A.Vinvp[lm];
    double ei[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) ei[p] = Ei[p] * vi;
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[p * 6 + q] -= ei[p] * Ej[q];
    if (diag && sl.x == sl.y) {
      const double gl = A.gp[lm];
#pragma unroll
      for (int p = 0; p < 6; ++p) bred[p] += ei[p] * gl;
    }
  }
  for (int e = A.L.st_ptr[blk] + lane, e1 = valid ? A.L.st_ptr[blk + 1] : 0; e < e1; e += G) {
    const int2 sl = A.L.st[e];
    const double* Ei = A.Et + (size_t)sl.x * 18;
    const double* Ej = A.Et + (size_t)sl.y * 18;
    const int lm = A.st_lm[sl.x];
    const double* Vi = A.Vinvt + (size_t)lm * 9;
    double ev[18];  // E_i V^-1 (6x3)
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int c = 0; c < 3; ++c) ev[p * 3 + c] = Ei[p * 3] * Vi[c] + Ei[p * 3 + 1] * Vi[3 + c] + Ei[p * 3 + 2] * Vi[6 + c];
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[p * 6 + q] -= ev[p * 3] * Ej[q * 3] + ev[p * 3 + 1] * Ej[q * 3 + 1] + ev[p * 3 + 2] * Ej[q * 3 + 2];
    if (diag && sl.x == sl.y) {
      const double* gl = A.gt + (size_t)lm * 3;
#pragma unroll
      for (int p = 0; p < 6; ++p) bred[p] += ev[p * 3] * gl[0] + ev[p * 3 + 1] * gl[1] + ev[p * 3 + 2] * gl[2];
    }
  }
  // ---- reduce over the G lanes of the group (butterfly: every lane ends with the total) ----
  constexpr int GW = G > 32 ? 32 : G;
#pragma unroll
  for (int k = 0; k < 36; ++k)
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (G >= 32 ? diag : __any_sync(0xffffffffu, diag)) {   // G >= 32: the branch is warp-uniform
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
      for (int o = GW / 2; o > 0; o >>= 1) {
        gr[k] += __shfl_xor_sync(0xffffffffu, gr[k], o);
        dd[k] += __shfl_xor_sync(0xffffffffu, dd[k], o);
        bred[k] += __shfl_xor_sync(0xffffffffu, bred[k], o);
      }
  }
  if (G == 128) {   // four warp totals -> one, in warp order; thread k (< 54) then owns value k
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < 36; ++k) xs[w * 54 + k] = acc[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) { xs[w * 54 + 36 + k] = gr[k]; xs[w * 54 + 42 + k] = dd[k]; xs[w * 54 + 48 + k] = bred[k]; }
    }
    __syncthreads();
    if (!valid || threadIdx.x >= 54) return;
    const int k = threadIdx.x;
    const double v = (xs[k] + xs[54 + k]) + (xs[108 + k] + xs[162 + k]);
    if (k < 36) A.Sblk[(size_t)blk * 36 + k] = v;
    if (diag) {   // gr, dd, bred of the same index live in threads 36+k, 42+k, 48+k: thread 36+k gathers its triple from smem
      if (k >= 36 && k < 42) {
        const int q = k - 36;
        const double g_ = v;
        const double d_ = (xs[42 + q] + xs[54 + 42 + q]) + (xs[108 + 42 + q] + xs[162 + 42 + q]);
        const double b_ = (xs[48 + q] + xs[54 + 48 + q]) + (xs[108 + 48 + q] + xs[162 + 48 + q]);
        A.udiag[6 * a + q] = d_; A.bvec[6 * a + q] = g_ - b_; A.graw[6 * a + q] = g_ / A.scale_c[6 * a + q];
      }
    }
    return;
  }
  if (!valid) return;
  double* out = A.Sblk + (size_t)blk * 36;
#pragma unroll
  for (int k = 0; k < 36; ++k)
    if ((k % G) == lane) out[k] = acc[k];
  if (diag) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (lane == k) { A.udiag[6 * a + k] = dd[k]; A.bvec[6 * a + k] = gr[k] - bred[k]; A.graw[6 * a + k] = gr[k] / sa[k]; }
  }
}
template <int G>
__global__ void __launch_bounds__(128) schur_block_kernel(BlockArgs A, const int* __restrict__ list, int nlist) {
  PDL_PROLOGUE();
  schur_block_body<G>(blockIdx.x, A, list, nlist);
}

// Diagonal blocks (a CTA each) and off-diagonal blocks (8 lanes each) in ONE launch: CTAs [0, n_diag) take the diagonal list.
// Measured and rejected in round 2 (profiles/r2_notes.md): 3 or 4 CTAs per SM through __launch_bounds__ (168 / 128 registers with
// spills: 55.3 / 68.4 us against 54.7), and an output-stationary variant (a lane per block entry, gather lists staged through
// shared memory with coalesced loads, no final shuffle reduction: 95 us — 7x the instructions per gather entry).
__global__ void __launch_bounds__(128) schur_merged_kernel(BlockArgs A, const int* __restrict__ diag_list, int n_diag,
                                                           const int* __restrict__ off_list, int n_off) {
  PDL_PROLOGUE();
  if ((int)blockIdx.x < n_diag) schur_block_body<128>(blockIdx.x, A, diag_list, n_diag);
  else schur_block_body<8>(blockIdx.x - n_diag, A, off_list, n_off);
}

// The LM damping term clamp(diag(J'J))/radius must see the GLOBAL diagonal: ranks exchange un-damped
// blocks plus udiag, and the damping is added while scattering into the dense matrix.
struct ScatterArgs {
  int nblk; const int* blk_a; const int* blk_b; const double* Sblk; const double* bvec; const double* udiag; double inv_radius;
  double* A; int ld; int n; int rows_total; int Rb;
  const int* doff;   // first column of each camera slot in the tile-aligned layout (nd_layout.h); padding columns keep the
                     // unit diagonal zero_tiles_kernel wrote
};
__global__ void scatter_kernel(ScatterArgs S) {
  PDL_PROLOGUE();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nb36 = S.nblk * 36;
  if (t < nb36) {
    const int blk = t / 36, k = t - blk * 36, p = k / 6, q = k - p * 6;
    const int a = S.blk_a[blk], b = S.blk_b[blk];
    double v = S.Sblk[t];
    if (a == b && p == q) v += lm_damp(S.udiag[6 * a + p], S.inv_radius);
    // block (a,b), a <= b, holds rows a / cols b; the lower triangle wants rows b / cols a
    const int R = S.doff[b] + q, Cc = S.doff[a] + p;
    if (a != b || R >= Cc) S.A[(size_t)R * S.ld + Cc] = v;
  } else {
    const int u = t - nb36;
    if (u < S.n) S.A[(size_t)S.Rb * S.ld + S.doff[u / 6] + u % 6] = S.bvec[u];   // b row
  }
}

// ---- landmark back-substitution, steps and candidate parameters ---------------------------------------
// cameras: delta_c = -y_c * scale ; x_c = Plus(x, delta); partial norms (one block, cams are few)
__device__ __forceinline__ void candidate_cams_body(const unsigned bid, int n_cams, const int* __restrict__ camslot, const int* __restrict__ doff,
                                                             const double* __restrict__ x, const double* __restrict__ yc, const double* __restrict__ scale_c,
                                                             double* __restrict__ delta_c, double* __restrict__ xc,
                                                             double* __restrict__ sc, double norm_weight) {
  __shared__ double sred[256];
  double step2 = 0.0, cn2 = 0.0;
  for (int k = threadIdx.x; k < n_cams; k += 256) {
    const int s = camslot[k];
    const double* xk = x + 7 * (size_t)k;
    double* ok = xc + 7 * (size_t)k;
    if (s < 0) {
#pragma unroll
      for (int c = 0; c < 7; ++c) ok[c] = xk[c];
      continue;
    }
    double d[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { d[c] = -yc[doff[s] + c] * scale_c[6 * s + c]; delta_c[6 * s + c] = d[c]; }
    // ceres::QuaternionParameterization::Plus
    const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double q[4];
    if (nrm > 0.0) {
      const double sn = sin(nrm) / nrm;
      const double z0 = cos(nrm), z1 = sn * d[0], z2 = sn * d[1], z3 = sn * d[2];
      q[0] = z0 * xk[0] - z1 * xk[1] - z2 * xk[2] - z3 * xk[3];
      q[1] = z0 * xk[1] + z1 * xk[0] + z2 * xk[3] - z3 * xk[2];
      q[2] = z0 * xk[2] - z1 * xk[3] + z2 * xk[0] + z3 * xk[1];
      q[3] = z0 * xk[3] + z1 * xk[2] - z2 * xk[1] + z3 * xk[0];
    } else {
      q[0] = xk[0]; q[1] = xk[1]; q[2] = xk[2]; q[3] = xk[3];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { ok[c] = q[c]; const double df = xk[c] - q[c]; step2 += df * df; cn2 += q[c] * q[c]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) { const double v = xk[4 + c] + d[3 + c]; ok[4 + c] = v; step2 += d[3 + c] * d[3 + c]; cn2 += v * v; }
  }
  const double a = block_sum_256(step2, sred);
  __syncthreads();
  const double b = block_sum_256(cn2, sred);
  if (threadIdx.x == 0) { sc[SC_STEP2] = a * norm_weight; sc[SC_CNORM2] = b * norm_weight; }
}
// squared norm of the free ambient parameters (x_norm at start)
__global__ void __launch_bounds__(256) xnorm_cams_kernel(int n_cams, const int* __restrict__ camslot, const double* __restrict__ x, double* out, double w) {
  PDL_PROLOGUE();
  __shared__ double sred[256];
  double a = 0.0;
  for (int k = threadIdx.x; k < n_cams; k += 256)
    if (camslot[k] >= 0)
      for (int c = 0; c < 7; ++c) a += x[7 * (size_t)k + c] * x[7 * (size_t)k + c];
  const double t = block_sum_256(a, sred);
  if (threadIdx.x == 0) *out = t * w;
}
template <int D>
__global__ void __launch_bounds__(256) xnorm_lm_kernel(int nv, const int* __restrict__ v_gl, const double* __restrict__ x, double* parts) {
  PDL_PROLOGUE();
  __shared__ double sred[256];
  const int v = blockIdx.x * 256 + threadIdx.x;
  double a = 0.0;
  if (v < nv)
    for (int k = 0; k < D; ++k) { const double t = x[(size_t)v_gl[v] * D + k]; a += t * t; }
  const double t = block_sum_256(a, sred);
  if (threadIdx.x == 0) parts[blockIdx.x] = t;
}

// ---- model cost change: -(J d)'(r + J d / 2) per observation ---------------------------------------------
template <int D, int ROWS, int JC>
__global__ void __launch_bounds__(256) model_cost_kernel(int n, const int* __restrict__ cs, const int* __restrict__ hs, const int* __restrict__ ls,
                                                         const double* __restrict__ J, const double* __restrict__ r,
                                                         const double* __restrict__ delta_c, const double* __restrict__ delta_l,
                                                         double* __restrict__ parts) {
  PDL_PROLOGUE();
  __shared__ double sred[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double acc = 0.0;
  if (i < n) {
    const int c = cs[i], h = hs[i], l = ls[i];
    double dc[6], dh[6], dl[D];
#pragma unroll
    for (int k = 0; k < 6; ++k) { dc[k] = c >= 0 ? delta_c[6 * c + k] : 0.0; dh[k] = h >= 0 ? delta_c[6 * h + k] : 0.0; }
#pragma unroll
    for (int k = 0; k < D; ++k) dl[k] = l >= 0 ? delta_l[l * D + k] : 0.0;
    const double* Ji = J + (size_t)i * ROWS * JC;
#pragma unroll
    for (int row = 0; row < ROWS; ++row) {
      double m = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) m += Ji[row * JC + k] * dc[k] + Ji[row * JC + 6 + k] * dh[k];
#pragma unroll
      for (int k = 0; k < D; ++k) m += Ji[row * JC + 12 + k] * dl[k];
      acc -= m * (r[(size_t)i * ROWS + row] + m * 0.5);
    }
  }
  const double t = block_sum_256(acc, sred);
  if (threadIdx.x == 0) parts[blockIdx.x] = t;
}

// ---- gradient max norm ||x - Plus(x, -g)||_inf (Ceres) -------------------------------------------------------
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void gmax_cams_body(const unsigned bid, int n_cams, const int* __restrict__ camslot, const double* __restrict__ x, const double* __restrict__ graw,
                                 double* __restrict__ mx) {
  const int k = bid * blockDim.x + threadIdx.x;
  if (k >= n_cams) return;
  const int s = camslot[k];
  if (s < 0) return;
  const double* xk = x + 7 * (size_t)k;
  const double d[3] = {-graw[6 * s], -graw[6 * s + 1], -graw[6 * s + 2]};
  const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double m = 0.0;
  if (nrm > 0.0) {
    const double sn = sin(nrm) / nrm;
    const double z0 = cos(nrm), z1 = sn * d[0], z2 = sn * d[1], z3 = sn * d[2];
    const double q0 = z0 * xk[0] - z1 * xk[1] - z2 * xk[2] - z3 * xk[3];
    const double q1 = z0 * xk[1] + z1 * xk[0] + z2 * xk[3] - z3 * xk[2];
    const double q2 = z0 * xk[2] - z1 * xk[3] + z2 * xk[0] + z3 * xk[1];
    const double q3 = z0 * xk[3] + z1 * xk[2] - z2 * xk[1] + z3 * xk[0];
    m = fmax(fmax(fabs(xk[0] - q0), fabs(xk[1] - q1)), fmax(fabs(xk[2] - q2), fabs(xk[3] - q3)));
  }
#pragma unroll
  for (int c = 3; c < 6; ++c) m = fmax(m, fabs(graw[6 * s + c]));
  atomic_max_nonneg(mx + MX_GMAX, m);
}
__device__ __forceinline__ void gmax_lm_body(const unsigned bid, int n, const double* __restrict__ g_scaled, const double* __restrict__ scale, double* __restrict__ mx) {
  const int i = bid * blockDim.x + threadIdx.x;
  // max over the bit patterns (what the atomic does; keeps a NaN visible), one atomic per warp instead of per landmark
  unsigned long long m = i < n ? (unsigned long long)__double_as_longlong(fabs(g_scaled[i] / scale[i])) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(mx + MX_GMAX), m);
}
// ---- everything between the reduced-system solve and the candidate evaluation, in ONE launch ---------------------------
// CTA ranges: landmark back-substitution fused with the candidate landmark values and their norm partials (points, then
// planes), the candidate cameras (one CTA), and the gradient max-norm pieces (cameras, points, planes) — six launches before.
template <int D>
__device__ __forceinline__ void backsub_candidate_body(const unsigned bid, int nv, const int* __restrict__ slot_ptr, const int* __restrict__ slot_cam,
                                                       const double* __restrict__ E, const double* __restrict__ Vinv, const double* __restrict__ g,
                                                       const double* __restrict__ yc, const int* __restrict__ doff, const double* __restrict__ scale_l,
                                                       double* __restrict__ delta_l, const int* __restrict__ v_gl, const double* __restrict__ x,
                                                       double* __restrict__ xc, double* __restrict__ parts) {
  __shared__ double sred[256];
  const int v = bid * 256 + threadIdx.x;
  double step2 = 0.0, cn2 = 0.0;
  if (v < nv) {
    double t[D];
#pragma unroll
    for (int a = 0; a < D; ++a) t[a] = g[v * D + a];
    for (int s = slot_ptr[v]; s < slot_ptr[v + 1]; ++s) {
      const double* y = yc + doff[slot_cam[s]];
      const double* Es = E + (size_t)s * 6 * D;
#pragma unroll
      for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int a = 0; a < D; ++a) t[a] -= Es[c * D + a] * y[c];
    }
    const int gidx = v_gl[v];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double y = 0.0;
#pragma unroll
      for (int b = 0; b < D; ++b) y += Vinv[(size_t)v * D * D + a * D + b] * t[b];
      const double d = -y * scale_l[v * D + a];
      delta_l[v * D + a] = d;
      const double val = x[(size_t)gidx * D + a] + d;
      xc[(size_t)gidx * D + a] = val;
      step2 += d * d; cn2 += val * val;
    }
  }
  const double a = block_sum_256(step2, sred);
  __syncthreads();
  const double b = block_sum_256(cn2, sred);
  if (threadIdx.x == 0) { parts[2 * bid] = a; parts[2 * bid + 1] = b; }
}

struct PostArgs {
  int nvp; const int* sp_ptr; const int* sp_cam; const double* Ep; const double* Vinvp; const double* gp; const double* scale_vp; double* delta_vp;
  const int* vp_gl; const double* x_rho; double* c_rho;
  int nvt; const int* st_ptr; const int* st_cam; const double* Et; const double* Vinvt; const double* gt; const double* scale_vt; double* delta_vt;
  const int* vt_gl; const double* x_theta; double* c_theta;
  const double* yc; const int* doff; double* parts_step;
  int K; const int* camslot; const double* x_cams; const double* scale_c; double* delta_c; double* c_cams; double* sc; double norm_weight;
  const double* graw; double* mx;
  int b_p, b_t, b_cam, b_gc, b_gp;   // CTA range ends: points | planes | candidate cameras | gmax cameras | gmax points | (rest) gmax planes
};
__global__ void __launch_bounds__(256) post_solve_kernel(PostArgs a) {
  PDL_PROLOGUE();
  const int b = blockIdx.x;
  if (b < a.b_p)
    backsub_candidate_body<1>(b, a.nvp, a.sp_ptr, a.sp_cam, a.Ep, a.Vinvp, a.gp, a.yc, a.doff, a.scale_vp, a.delta_vp, a.vp_gl, a.x_rho, a.c_rho, a.parts_step);
  else if (b < a.b_t)
    backsub_candidate_body<3>(b - a.b_p, a.nvt, a.st_ptr, a.st_cam, a.Et, a.Vinvt, a.gt, a.yc, a.doff, a.scale_vt, a.delta_vt, a.vt_gl, a.x_theta, a.c_theta,
                              a.parts_step + 2 * a.b_p);
  else if (b < a.b_cam) candidate_cams_body(0, a.K, a.camslot, a.doff, a.x_cams, a.yc, a.scale_c, a.delta_c, a.c_cams, a.sc, a.norm_weight);
  else if (b < a.b_gc) gmax_cams_body(b - a.b_cam, a.K, a.camslot, a.x_cams, a.graw, a.mx);
  else if (b < a.b_gp) gmax_lm_body(b - a.b_gc, a.nvp, a.gp, a.scale_vp, a.mx);
  else gmax_lm_body(b - a.b_gp, 3 * a.nvt, a.gt, a.scale_vt, a.mx);
}

// ---- end of an LM iteration's device work --------------------------------------------------------------------------
// ONE launch finishes the three two-level sums of the iteration (step / candidate norms, model cost change, candidate cost;
// same summation order as sum_parts(_pair)_kernel), folds the factorisation's failure flag into mx[], and — single GPU —
// publishes the 10 scalars the host's accept / reject logic needs straight into page-locked host memory, followed by a
// sequence number the host spins on (no copy-engine round trip, no stream synchronisation). It also re-arms mx[] and the
// failure flag for the next iteration. Multi-GPU: host == nullptr here, the scalars are all-reduced first and
// publish_scalars_kernel does the publishing.
struct TailArgs {
  const double* parts_step; int n_step;   // interleaved (step^2, candidate-norm^2) partial pairs of post_solve_kernel's landmark CTAs
  const double* parts_mcc; int n_mcc;     // model_cost_kernel partials
  const double* parts_cand; int n_cand;   // interleaved (cost, fixed cost) partial pairs of the candidate evaluation
  double* sc; double* mx; int* fail;
  double* host; double seq;
  double* gather; int rank, world;        // multi-GPU: world x (SC_N + MX_N) slots, this rank fills its own and zeroes the others
};
constexpr int H_SEQ = 63;                 // slot of the sequence number in tslam_ctx::h_scalars (64 doubles)

__device__ __forceinline__ void publish_to_host(const double* sc, double* mx, int* fail, double* host, double seq) {
  // one thread, program order: 10 payload stores, system-scope fence, then the sequence number
  for (int k = 0; k < SC_N; ++k) host[k] = sc[k];
  for (int k = 0; k < MX_N; ++k) { host[SC_N + k] = mx[k]; mx[k] = 0.0; }
  *fail = 0;
  __threadfence_system();
  *reinterpret_cast<volatile double*>(host + H_SEQ) = seq;
}
__global__ void __launch_bounds__(256) iteration_tail_kernel(TailArgs a) {
  PDL_PROLOGUE();
  // five sums at once: per-thread strided partials (all loads independent), then ONE tree over the five columns — the same
  // pairing of additions as five sum_parts(_pair)_kernel launches, with 8 barriers instead of 45
  __shared__ double s[5][256];
  {
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
    for (int i = threadIdx.x; i < a.n_step; i += 256) { p0 += a.parts_step[2 * (size_t)i]; p1 += a.parts_step[2 * (size_t)i + 1]; }
    for (int i = threadIdx.x; i < a.n_mcc; i += 256) p2 += a.parts_mcc[i];
    for (int i = threadIdx.x; i < a.n_cand; i += 256) { p3 += a.parts_cand[2 * (size_t)i]; p4 += a.parts_cand[2 * (size_t)i + 1]; }
    s[0][threadIdx.x] = p0; s[1][threadIdx.x] = p1; s[2][threadIdx.x] = p2; s[3][threadIdx.x] = p3; s[4][threadIdx.x] = p4;
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int k = 0; k < 5; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
    }
    __syncthreads();
  }
  const double step2 = s[0][0], cn2 = s[1][0], mcc = s[2][0], cand = s[3][0], cand_fixed = s[4][0];
  if (threadIdx.x == 0) {
    a.sc[SC_STEP2] += step2; a.sc[SC_CNORM2] += cn2;   // the candidate-camera CTA of post_solve_kernel wrote the camera part
    a.sc[SC_MCC] = mcc; a.sc[SC_CAND] = cand; a.sc[SC_CAND_FIXED] = cand_fixed;
    if (*a.fail) a.mx[MX_FAIL] = fmax(a.mx[MX_FAIL], 1.0);
    if (a.host) publish_to_host(a.sc, a.mx, a.fail, a.host, a.seq);
    if (a.gather)
      for (int r = 0; r < a.world; ++r) {
        double* slot = a.gather + (size_t)r * (SC_N + MX_N);
        for (int k = 0; k < SC_N; ++k) slot[k] = r == a.rank ? a.sc[k] : 0.0;
        for (int k = 0; k < MX_N; ++k) slot[SC_N + k] = r == a.rank ? a.mx[k] : 0.0;
      }
  }
}
// Multi-GPU: `gather` has been summed across the ranks (every slot was non-zero on exactly one rank, so the sum is an
// all-gather): sums in rank order and maxima over the ranks — identical on every rank — then the publish as on one GPU.
__global__ void publish_gathered_kernel(const double* gather, int world, double* mx, int* fail, double* host, double seq) {
  PDL_PROLOGUE();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int k = 0; k < SC_N; ++k) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += gather[(size_t)r * (SC_N + MX_N) + k];
    host[k] = s;
  }
  for (int k = 0; k < MX_N; ++k) {
    unsigned long long m = 0ull;   // non-negative doubles: order of the bit patterns (keeps a NaN visible, like the atomics that fill mx[])
    for (int r = 0; r < world; ++r) { const unsigned long long v = (unsigned long long)__double_as_longlong(gather[(size_t)r * (SC_N + MX_N) + SC_N + k]); m = v > m ? v : m; }
    host[SC_N + k] = __longlong_as_double((long long)m);
    mx[k] = 0.0;
  }
  *fail = 0;
  __threadfence_system();
  *reinterpret_cast<volatile double*>(host + H_SEQ) = seq;
}
// final residual scatter into the global residual vector (multi-GPU: other ranks' entries stay 0)
__global__ void scatter_rows_kernel(int n, int width, const int* __restrict__ gsel, const double* __restrict__ src, double* __restrict__ dst) {
  PDL_PROLOGUE();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const int i = t / width, k = t - i * width;
  dst[(size_t)(gsel ? gsel[i] : i) * width + k] = src[t];
}
template <int D>
__global__ void export_lm_kernel(int nv, const int* __restrict__ v_gl, const double* __restrict__ x, double* __restrict__ out, double* __restrict__ owners) {
  PDL_PROLOGUE();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nv * D) return;
  const int v = t / D, a = t - v * D;
  out[(size_t)v_gl[v] * D + a] = x[(size_t)v_gl[v] * D + a];
  if (a == 0) owners[v_gl[v]] = 1.0;
}
// after the all-reduce: landmarks somebody owns take the reduced value, all others keep what the caller passed in
template <int D>
__global__ void merge_lm_kernel(int n, const double* __restrict__ reduced, const double* __restrict__ owners, double* __restrict__ x) {
  PDL_PROLOGUE();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * D) return;
  if (owners[t / D] > 0.0) x[t] = reduced[t];
}

// ================================================================================================
// host side
// ================================================================================================
struct Solver : SolverIndex {
  tslam_ctx* ctx = nullptr;
  tslam_dev_problem* d = nullptr;
  // values
  DevBuf<double> xc_cams, xc_rho, xc_theta;
  DevBuf<double> pr, pJ, tr, tJ, cr_p, cr_t, pJ2, tJ2;   // two (residual, Jacobian) sets: linearisation point and candidate
  DevBuf<double> tG;                                     // Gram matrices of the text blocks at the linearisation point
  double *rp = nullptr, *Jp = nullptr, *rt = nullptr, *Jt = nullptr;        // current linearisation point
  double *rp2 = nullptr, *Jp2 = nullptr, *rt2 = nullptr, *Jt2 = nullptr;    // candidate (swapped in when a step is accepted)
  bool spec_J = false;   // evaluate the Jacobian together with the candidate cost (saves the re-evaluation after an accepted step)
  DevBuf<double> scale_c, scale_vp, scale_vt, colnorm_c;
  DevBuf<double> Vp, gp, Vinvp, Vt, gt, Vinvt, Ep, Et;
  DevBuf<double> red;        // [Sblk | b | graw | udiag]  (one all-reduce)
  DevBuf<double> scl, scr;   // local scalars / reduced copy
  DevBuf<double> mx;         // max-reduced scalars
  DevBuf<double> A, ywork, yc, delta_c, delta_vp, delta_vt, parts, parts_step, parts_mcc, gat;
  int tail_n_cand = 0;   // partial pairs of the deferred candidate-cost sum (iteration_tail_kernel)
  DevBuf<int> fail;
  size_t red_n = 0;
  double *Sblk = nullptr, *bvec = nullptr, *graw = nullptr, *udiag = nullptr, *sc = nullptr;
  double* x_cams = nullptr; double* x_rho = nullptr; double* x_theta = nullptr;   // current (alias d-> buffers or xc)
  double* c_cams = nullptr; double* c_rho = nullptr; double* c_theta = nullptr;   // candidate
  double setup_ms = 0;
  // phase timing
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_kind;   // phase that STARTS at this mark (0..6), 7 = end marker
  int ev_used = 0;
  ~Solver() { for (auto e : ev) cudaEventDestroy(e); }
};

template <typename T, typename Vec>
static cudaError_t up(DevBuf<T>& b, const Vec& h, cudaStream_t s) { return b.upload(h.data(), h.size(), s); }

template <typename T, typename Vec>
static cudaError_t up_as(DevBuf<T>& b, const Vec& h, cudaStream_t s) {
  static_assert(sizeof(T) == sizeof(typename Vec::value_type), "layout-compatible element types expected");
  return b.upload(reinterpret_cast<const T*>(h.data()), h.size(), s);
}

// Host path (sharded / multi-GPU problems, or TSLAM_HOST_ANALYSIS=1): analysis.cpp + upload of the index structures.
static int analyze_on_host_and_upload(Solver& S, Analysis& A, std::chrono::steady_clock::time_point& T1) {
  tslam_ctx* ctx = S.ctx; tslam_dev_problem* d = S.d;
  cudaStream_t st = ctx->stream;
  if (!d->have_host_index) return set_error(TSLAM_ERR_ARG, "host structure analysis requested but the index copies were not kept");
  IndexView V;
  V.n_cams = d->n_cams; V.n_points = d->n_points; V.n_planes = d->n_planes; V.g_pobs = d->g_pobs; V.g_tobs = d->g_tobs;
  V.cam_fixed = d->h_cam_fixed.data(); V.rho_fixed = d->h_rho_fixed.data(); V.theta_fixed = d->h_theta_fixed.data();
  V.p_cam = d->h_p_cam.data(); V.p_host = d->h_p_host.data(); V.p_lm = d->h_p_lm.data();
  V.t_cam = d->h_t_cam.data(); V.t_host = d->h_t_host.data(); V.t_plane = d->h_t_plane.data();
  V.lp = d->n_pobs; V.lt = d->n_tobs;
  V.gsel_p = d->sharded ? d->gsel_p.data() : nullptr; V.gsel_t = d->sharded ? d->gsel_t.data() : nullptr;
  if (!ctx->host_arena)
    ctx->host_arena = new Arena([](size_t n) -> void* { void* q = nullptr; return cudaHostAlloc(&q, n, cudaHostAllocDefault) == cudaSuccess ? q : nullptr; },
                                [](void* q) { cudaFreeHost(q); });
  try { analyze_structure(V, A, *ctx->host_arena); } catch (const std::exception& e) { return set_error(TSLAM_ERR_CUDA, "structure analysis failed: %s", e.what()); }
  S.K = A.K; S.nc = A.nc; S.nl = A.nl; S.npl = A.npl; S.lp = A.lp; S.lt = A.lt;
  S.nvp = A.nvp; S.nvt = A.nvt; S.nsp = A.nsp; S.nst = A.nst; S.nblk = A.nblk;
  S.n = A.n; S.npad = A.npad; S.ld = A.ld; S.rows = A.rows; S.Tn = A.Tn;
  T1 = std::chrono::steady_clock::now();
  // ---- upload ----
  int rc = chol_upload(ctx, A.chol, &S.chol);
  if (rc) return rc;
  TSL_CUDA(up(S.camslot_d, A.camslot, st)); TSL_CUDA(up(S.doff, A.doff, st));
  TSL_CUDA(up(S.p_cs, A.p_cs, st)); TSL_CUDA(up(S.p_hs, A.p_hs, st)); TSL_CUDA(up(S.p_ls, A.LP.obs_ls, st));
  TSL_CUDA(up(S.t_cs, A.t_cs, st)); TSL_CUDA(up(S.t_hs, A.t_hs, st)); TSL_CUDA(up(S.t_ls, A.LT.obs_ls, st));
  TSL_CUDA(up(S.p_active, A.p_act, st)); TSL_CUDA(up(S.t_active, A.t_act, st)); TSL_CUDA(up(S.t_fmask, A.t_fm, st));
  TSL_CUDA(up(S.vp_gl, A.LP.v_gl, st)); TSL_CUDA(up(S.vt_gl, A.LT.v_gl, st));
  TSL_CUDA(up(S.vp_obs_ptr, A.LP.obs_ptr, st)); TSL_CUDA(up(S.vp_obs, A.LP.obs, st));
  TSL_CUDA(up(S.vt_obs_ptr, A.LT.obs_ptr, st)); TSL_CUDA(up(S.vt_obs, A.LT.obs, st));
  TSL_CUDA(up(S.sp_ptr, A.LP.slot_ptr, st)); TSL_CUDA(up(S.sp_cam, A.LP.slot_cam, st)); TSL_CUDA(up(S.sp_lm, A.LP.slot_lm, st));
  TSL_CUDA(up(S.spe_ptr, A.LP.ent_ptr, st)); TSL_CUDA(up(S.spe, A.LP.ent, st));
  TSL_CUDA(up(S.st_ptr, A.LT.slot_ptr, st)); TSL_CUDA(up(S.st_cam, A.LT.slot_cam, st)); TSL_CUDA(up(S.st_lm, A.LT.slot_lm, st));
  TSL_CUDA(up(S.ste_ptr, A.LT.ent_ptr, st)); TSL_CUDA(up(S.ste, A.LT.ent, st));
  TSL_CUDA(up(S.blk_a, A.blk_a, st)); TSL_CUDA(up(S.blk_b, A.blk_b, st)); TSL_CUDA(up(S.diag_blk, A.diag_blk, st));
  TSL_CUDA(up(S.offdiag_blk, A.offdiag_blk, st)); S.noff = (int)A.offdiag_blk.size();
  S.est_entries = (long long)A.bdp.size() + (long long)A.bdt.size() + (long long)A.bsp.size() + (long long)A.bst.size();
  TSL_CUDA(up(S.bdp_ptr, A.bdp_ptr, st)); TSL_CUDA(up(S.bdp, A.bdp, st)); TSL_CUDA(up(S.bdt_ptr, A.bdt_ptr, st)); TSL_CUDA(up(S.bdt, A.bdt, st));
  TSL_CUDA(up(S.bsp_ptr, A.bsp_ptr, st)); TSL_CUDA(up_as(S.bsp, A.bsp, st)); TSL_CUDA(up(S.bst_ptr, A.bst_ptr, st)); TSL_CUDA(up_as(S.bst, A.bst, st));
  if (d->sharded) { TSL_CUDA(up(S.gsel_p, d->gsel_p, st)); TSL_CUDA(up(S.gsel_t, d->gsel_t, st)); }
  return TSLAM_OK;
}

// Structure analysis (device or host) + allocation of the value buffers.
static int analyze_and_upload(Solver& S) {
  auto T0 = std::chrono::steady_clock::now();
  static const bool trace_setup = getenv("TSLAM_SETUP_TRACE") != nullptr;
  tslam_ctx* ctx = S.ctx; tslam_dev_problem* d = S.d;
  cudaStream_t st = ctx->stream;
  double dev_laps[4] = {0, 0, 0, 0};
  const bool on_device = device_analysis_supported(ctx, d);
  Analysis A;
  auto T1 = T0;
  if (on_device) {
    // unsharded problem: the index structures are built where the observation arrays already are (analysis_dev.cu)
    int rc = analyze_structure_device(ctx, d, S, dev_laps);
    if (rc) return rc;
    T1 = std::chrono::steady_clock::now();
  } else {
    int rc = analyze_on_host_and_upload(S, A, T1);
    if (rc) return rc;
  }
  const int K = S.K, nc = S.nc, lp = S.lp, lt = S.lt;
  // ---- value buffers ----
  TSL_CUDA(S.xc_cams.reserve(7 * (size_t)K)); TSL_CUDA(S.xc_rho.reserve(d->n_points)); TSL_CUDA(S.xc_theta.reserve(3 * (size_t)d->n_planes));
  TSL_CUDA(S.pr.reserve(2 * (size_t)lp)); TSL_CUDA(S.pJ.reserve(26 * (size_t)lp)); TSL_CUDA(S.cr_p.reserve(2 * (size_t)lp)); TSL_CUDA(S.pJ2.reserve(26 * (size_t)lp));
  TSL_CUDA(S.tr.reserve(8 * (size_t)lt)); TSL_CUDA(S.tJ.reserve(120 * (size_t)lt)); TSL_CUDA(S.cr_t.reserve(8 * (size_t)lt)); TSL_CUDA(S.tJ2.reserve(120 * (size_t)lt)); TSL_CUDA(S.tG.reserve((size_t)TG_STRIDE * lt));
  TSL_CUDA(S.scale_c.reserve(6 * (size_t)nc)); TSL_CUDA(S.colnorm_c.reserve(6 * (size_t)nc));
  TSL_CUDA(S.scale_vp.reserve(S.nvp)); TSL_CUDA(S.scale_vt.reserve(3 * (size_t)S.nvt));
  TSL_CUDA(S.Vp.reserve(S.nvp)); TSL_CUDA(S.gp.reserve(S.nvp)); TSL_CUDA(S.Vinvp.reserve(S.nvp));
  TSL_CUDA(S.Vt.reserve(9 * (size_t)S.nvt)); TSL_CUDA(S.gt.reserve(3 * (size_t)S.nvt)); TSL_CUDA(S.Vinvt.reserve(9 * (size_t)S.nvt));
  TSL_CUDA(S.Ep.reserve(6 * (size_t)S.nsp)); TSL_CUDA(S.Et.reserve(18 * (size_t)S.nst));
  S.red_n = (size_t)S.nblk * 36 + 18 * (size_t)nc;
  TSL_CUDA(S.red.reserve(S.red_n));
  S.Sblk = S.red.p; S.bvec = S.red.p + (size_t)S.nblk * 36; S.graw = S.bvec + 6 * (size_t)nc; S.udiag = S.graw + 6 * (size_t)nc;
  TSL_CUDA(cudaMemsetAsync(S.red.p, 0, (S.red_n ? S.red_n : 1) * sizeof(double), st));
  TSL_CUDA(S.scl.reserve(SC_N)); TSL_CUDA(S.scr.reserve(SC_N));
  S.sc = S.scl.p;
  TSL_CUDA(cudaMemsetAsync(S.scl.p, 0, SC_N * sizeof(double), st));
  TSL_CUDA(S.mx.reserve(MX_N));
  TSL_CUDA(S.A.reserve((size_t)S.rows * S.ld)); TSL_CUDA(S.ywork.reserve(S.ld)); TSL_CUDA(S.yc.reserve(S.ld));
  TSL_CUDA(S.delta_c.reserve(6 * (size_t)nc)); TSL_CUDA(S.delta_vp.reserve(S.nvp)); TSL_CUDA(S.delta_vt.reserve(3 * (size_t)S.nvt));
  const size_t nparts = 2 * ((size_t)(lp + 127) / 128 + (size_t)(8 * (size_t)lt + 127) / 128) + 2 * ((size_t)(S.nvp + 255) / 256 + (size_t)(S.nvt + 255) / 256) + 64;
  TSL_CUDA(S.parts.reserve(nparts));
  TSL_CUDA(S.parts_step.reserve(2 * ((size_t)(S.nvp + 255) / 256 + (size_t)(S.nvt + 255) / 256) + 2));
  TSL_CUDA(S.parts_mcc.reserve((size_t)(lp + 255) / 256 + (size_t)(lt + 255) / 256 + 2));
  TSL_CUDA(S.fail.reserve(1));
  TSL_CUDA(S.gat.reserve((size_t)(ctx->world > 1 ? ctx->world : 1) * (SC_N + MX_N)));
  TSL_CUDA(cudaStreamSynchronize(st));   // the host vectors of A go out of scope on return
  // host copies the solver keeps (the arena is recycled by the next analysis on this context)
  S.camslot.assign(A.camslot.begin(), A.camslot.end()); S.vp_gl_h.assign(A.LP.v_gl.begin(), A.LP.v_gl.end()); S.vt_gl_h.assign(A.LT.v_gl.begin(), A.LT.v_gl.end());
  S.lmfree_p_h.assign(A.lmfree_p.begin(), A.lmfree_p.end()); S.lmfree_t_h.assign(A.lmfree_t.begin(), A.lmfree_t.end());
  auto T2 = std::chrono::steady_clock::now();
  S.setup_ms = std::chrono::duration<double, std::milli>(T2 - T0).count();
  if (trace_setup && on_device)
    fprintf(stderr, "[tslam setup] device analysis %.3f ms (layout+slots+block table+round trip %.3f, symbolic+upload %.3f, block arrays+gather lists %.3f); "
            "value buffers %.3f ms\n", dev_laps[3], dev_laps[0], dev_laps[1], dev_laps[2], std::chrono::duration<double, std::milli>(T2 - T1).count());
  else if (trace_setup)
    fprintf(stderr, "[tslam setup] host analysis %.3f ms (layout+order %.3f, landmark side %.3f, block structure %.3f, symbolic %.3f, gather lists %.3f); "
            "index upload + buffers %.3f ms\n", std::chrono::duration<double, std::milli>(T1 - T0).count(), A.lap_ms[0], A.lap_ms[1], A.lap_ms[2], A.lap_ms[3],
            A.lap_ms[4], std::chrono::duration<double, std::milli>(T2 - T1).count());
  return TSLAM_OK;
}

// ---- phase timing helpers -----------------------------------------------------------------------------
static void mark(Solver& S, int kind) {
  if (!S.timing) return;
  if (S.ev_used == (int)S.ev.size()) { cudaEvent_t e; cudaEventCreate(&e); S.ev.push_back(e); S.ev_kind.push_back(0); }
  S.ev_kind[S.ev_used] = kind;
  cudaEventRecord(S.ev[S.ev_used++], S.ctx->stream);
}

static inline int grid_for(int n, int b) { return n > 0 ? (n + b - 1) / b : 0; }

// Evaluate residuals (+ Jacobians) at (cams, rho, theta); cost -> sc[cost_slot], sc[cost_slot+1]
static int eval_at(Solver& S, const double* cams, const double* rho, const double* theta, bool want_J, int jac_mode, int cost_slot,
                   double* pr, double* tr, double* pJ, double* tJ, bool defer_sum = false) {
  tslam_ctx* ctx = S.ctx; tslam_dev_problem* d = S.d;
  int np = 0, nt = 0;
  double* parts = S.parts.p;
  int rc = launch_eval_points_robust(ctx, d, cams, rho, S.p_active.p, pr, want_J ? pJ : nullptr, parts, &np);
  if (rc) return rc;
  rc = launch_eval_text_robust(ctx, d, cams, theta, S.t_active.p, S.t_fmask.p, jac_mode, tr, want_J ? tJ : nullptr, parts + 2 * np, &nt);
  if (rc) return rc;
  S.tail_n_cand = np + nt;
  if (!defer_sum) LAUNCH(launch_k(sum_parts_pair_kernel, 2, 256, 0, ctx->stream, parts, np + nt, S.sc + cost_slot, S.sc + cost_slot + 1, 0));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

static BlockLists block_lists(Solver& S) {
  BlockLists L;
  L.dp_ptr = S.bdp_ptr.p; L.dp = S.bdp.p; L.dt_ptr = S.bdt_ptr.p; L.dt = S.bdt.p;
  L.sp_ptr = S.bsp_ptr.p; L.sp = S.bsp.p; L.st_ptr = S.bst_ptr.p; L.st = S.bst.p;
  return L;
}

// after a new Jacobian: V, g, E in the scaled system
static int accumulate_landmarks(Solver& S) {
  cudaStream_t st = S.ctx->stream;
  if (S.nvp) {
    const int g1 = grid_for(S.nvp, 128), g2 = grid_for(S.nsp, 128);
    LAUNCH(launch_k(accum_merged_kernel<1, 2, 13, false>, g1 + g2, 128, 0, st, g1, S.nvp, S.vp_obs_ptr.p, S.vp_obs.p, S.Jp, S.rp, S.scale_vp.p, S.Vp.p, S.gp.p,
                    S.nsp, S.spe_ptr.p, S.spe.p, S.sp_cam.p, S.sp_lm.p, S.scale_c.p, S.Ep.p));
  }
  if (S.lt) LAUNCH(launch_k(text_gram_kernel, grid_for(S.lt, 4), 128, 0, st, S.lt, S.Jt, S.rt, S.tG.p));
  if (S.nvt) {
    const int g1 = grid_for(S.nvt * 32, 128), g2 = grid_for(S.nst * 32, 128);
    LAUNCH(launch_k(accum_merged_kernel<3, 8, 15, true>, g1 + g2, 128, 0, st, g1, S.nvt, S.vt_obs_ptr.p, S.vt_obs.p, S.Jt, S.rt, S.scale_vt.p, S.Vt.p, S.gt.p,
                    S.nst, S.ste_ptr.p, S.ste.p, S.st_cam.p, S.st_lm.p, S.scale_c.p, S.Et.p));
  }
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

static int compute_jacobi_scaling(Solver& S) {
  cudaStream_t st = S.ctx->stream;
  const int nc = S.nc;
  // unscaled column norms: cameras via the diagonal blocks' direct entries, landmarks via V with scale == 1
  if (nc) {
    LAUNCH(launch_k(cam_colnorm_kernel, grid_for(nc * 32, 128), 128, 0, st, nc, S.diag_blk.p, block_lists(S), S.Jp, S.Jt, S.colnorm_c.p));
    TSL_CHECK_LAUNCH();
    int rc = comm_allreduce_sum(S.ctx, S.colnorm_c.p, 6 * (size_t)nc);
    if (rc) return rc;
    LAUNCH(launch_k(scale_from_norm_kernel, grid_for(6 * nc, 256), 256, 0, st, 6 * nc, S.colnorm_c.p, S.scale_c.p));
  }
  if (S.nvp) {
    LAUNCH(launch_k(fill_kernel, grid_for(S.nvp, 256), 256, 0, st, S.scale_vp.p, S.nvp, 1.0));
    LAUNCH(launch_k(lm_accum_kernel<1, 2, 13>, grid_for(S.nvp, 128), 128, 0, st, S.nvp, S.vp_obs_ptr.p, S.vp_obs.p, S.Jp, S.rp, S.scale_vp.p, S.Vp.p, S.gp.p));
    LAUNCH(launch_k(lm_scale_kernel<1>, grid_for(S.nvp, 256), 256, 0, st, S.nvp, S.Vp.p, S.scale_vp.p));
  }
  if (S.nvt) {
    LAUNCH(launch_k(fill_kernel, grid_for(3 * S.nvt, 256), 256, 0, st, S.scale_vt.p, 3 * S.nvt, 1.0));
    LAUNCH(launch_k(lm_accum_warp_kernel<3, 8, 15>, grid_for(S.nvt * 32, 128), 128, 0, st, S.nvt, S.vt_obs_ptr.p, S.vt_obs.p, S.Jt, S.rt, S.scale_vt.p, S.Vt.p, S.gt.p));
    LAUNCH(launch_k(lm_scale_kernel<3>, grid_for(3 * S.nvt, 256), 256, 0, st, S.nvt, S.Vt.p, S.scale_vt.p));
  }
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

// One linear solve for the current radius: builds the reduced system, reduces it across ranks,
// factors it and back-substitutes; leaves delta_* and the candidate parameters on the device.
static int compute_step(Solver& S, double radius) {
  tslam_ctx* ctx = S.ctx; cudaStream_t st = ctx->stream;
  const double inv_radius = 1.0 / radius;
  const int nc = S.nc;
  mark(S, 1);  // landmark / Schur prep
  // mx[] and the failure flag are zero here: cleared by run_lm before the first iteration, re-armed by every publish
  if (S.nvp) LAUNCH(launch_k(lm_vinv_kernel<1>, grid_for(S.nvp, 256), 256, 0, st, S.nvp, S.Vp.p, inv_radius, S.Vinvp.p, S.mx.p));
  if (S.nvt) LAUNCH(launch_k(lm_vinv_kernel<3>, grid_for(S.nvt, 128), 128, 0, st, S.nvt, S.Vt.p, inv_radius, S.Vinvt.p, S.mx.p));
  mark(S, 2);  // reduced system build
  if (S.nblk) {
    BlockArgs B;
    B.nblk = S.nblk; B.blk_a = S.blk_a.p; B.blk_b = S.blk_b.p; B.L = block_lists(S);
    B.pJ = S.Jp; B.pr = S.rp; B.tJ = S.Jt; B.tr = S.rt; B.tG = S.tG.p; B.scale_c = S.scale_c.p;
    B.Ep = S.Ep.p; B.Vinvp = S.Vinvp.p; B.gp = S.gp.p; B.sp_lm = S.sp_lm.p;
    B.Et = S.Et.p; B.Vinvt = S.Vinvt.p; B.gt = S.gt.p; B.st_lm = S.st_lm.p;
    B.Sblk = S.Sblk; B.bvec = S.bvec; B.graw = S.graw; B.udiag = S.udiag;
    // diagonal blocks (heavy gather lists, ~650 entries on the global-BA shape) get a CTA each, off-diagonal blocks 8 lanes —
    // unless the camera graph is small and dense (local BA: 7 cameras, 28 blocks, ~600 entries each): then every block gets a CTA
    if (S.est_entries > 96 * (long long)S.nblk) {
      LAUNCH(launch_k(schur_block_kernel<128>, S.nblk, 128, 0, st, B, nullptr, S.nblk));
    } else {
      LAUNCH(launch_k(schur_merged_kernel, S.nc + grid_for(S.noff * 8, 128), 128, 0, st, B, S.diag_blk.p, S.nc, S.offdiag_blk.p, S.noff));
    }
    TSL_CHECK_LAUNCH();
  }
  mark(S, 3);  // all-reduce
  if (ctx->world > 1) {
    int rc = comm_allreduce_sum(ctx, S.red.p, S.red_n);
    if (rc) return rc;
  }
  mark(S, 4);  // Cholesky
  if (nc) {
    { int rc = chol_clear(ctx, S.chol, S.A.p); if (rc) return rc; }
    ScatterArgs Sa{S.nblk, S.blk_a.p, S.blk_b.p, S.Sblk, S.bvec, S.udiag, inv_radius, S.A.p, S.ld, S.n, S.rows, S.Tn * 64, S.doff.p};
    const int total = S.nblk * 36 + S.n;
    LAUNCH(launch_k(scatter_kernel, grid_for(total, 256), 256, 0, st, Sa));
    TSL_CHECK_LAUNCH();
    int rc = chol_solve(ctx, S.chol, S.A.p, S.ywork.p, S.yc.p, S.fail.p);
    if (rc) return rc;
  }
  mark(S, 5);  // back-substitution + candidate
  {
    const int gvp = grid_for(S.nvp, 256), gvt = grid_for(S.nvt, 256);
    PostArgs a;
    a.nvp = S.nvp; a.sp_ptr = S.sp_ptr.p; a.sp_cam = S.sp_cam.p; a.Ep = S.Ep.p; a.Vinvp = S.Vinvp.p; a.gp = S.gp.p; a.scale_vp = S.scale_vp.p;
    a.delta_vp = S.delta_vp.p; a.vp_gl = S.vp_gl.p; a.x_rho = S.x_rho; a.c_rho = S.c_rho;
    a.nvt = S.nvt; a.st_ptr = S.st_ptr.p; a.st_cam = S.st_cam.p; a.Et = S.Et.p; a.Vinvt = S.Vinvt.p; a.gt = S.gt.p; a.scale_vt = S.scale_vt.p;
    a.delta_vt = S.delta_vt.p; a.vt_gl = S.vt_gl.p; a.x_theta = S.x_theta; a.c_theta = S.c_theta;
    a.yc = S.yc.p; a.doff = S.doff.p; a.parts_step = S.parts_step.p;   // pairs summed onto sc[SC_STEP2], sc[SC_CNORM2] by iteration_tail_kernel
    a.K = S.K; a.camslot = S.camslot_d.p; a.x_cams = S.x_cams; a.scale_c = S.scale_c.p; a.delta_c = S.delta_c.p; a.c_cams = S.c_cams; a.sc = S.sc;
    a.norm_weight = ctx->rank == 0 ? 1.0 : 0.0;
    a.graw = S.graw; a.mx = S.mx.p;   // graw (cams) comes from the reduced-system build and is already all-reduced; landmark gradients are local
    a.b_p = gvp; a.b_t = a.b_p + gvt; a.b_cam = a.b_t + 1; a.b_gc = a.b_cam + grid_for(S.K, 256); a.b_gp = a.b_gc + grid_for(S.nvp, 256);
    const int grid = a.b_gp + grid_for(3 * S.nvt, 256);
    LAUNCH(launch_k(post_solve_kernel, grid, 256, 0, st, a));
  }
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

static int model_and_candidate_cost(Solver& S, int jac_mode) {
  cudaStream_t st = S.ctx->stream;
  mark(S, 6);  // model cost change + candidate cost
  const int gp = grid_for(S.lp, 256), gt = grid_for(S.lt, 256);
  double* parts = S.parts_mcc.p;
  if (S.lp) LAUNCH(launch_k(model_cost_kernel<1, 2, 13>, gp, 256, 0, st, S.lp, S.p_cs.p, S.p_hs.p, S.p_ls.p, S.Jp, S.rp, S.delta_c.p, S.delta_vp.p, parts));
  if (S.lt) LAUNCH(launch_k(model_cost_kernel<3, 8, 15>, gt, 256, 0, st, S.lt, S.t_cs.p, S.t_hs.p, S.t_ls.p, S.Jt, S.rt, S.delta_c.p, S.delta_vt.p, parts + gp));
  TSL_CHECK_LAUNCH();
  int rc = eval_at(S, S.c_cams, S.c_rho, S.c_theta, S.spec_J, jac_mode, SC_CAND, S.rp2, S.rt2, S.Jp2, S.Jt2, /*defer_sum=*/true);
  if (rc) return rc;
  // the three sums of this iteration (+ the publish on one GPU) in one launch
  tslam_ctx* ctx = S.ctx;
  const int gvp = grid_for(S.nvp, 256), gvt = grid_for(S.nvt, 256);
  TailArgs ta{S.parts_step.p, gvp + gvt, S.parts_mcc.p, gp + gt, S.parts.p, S.tail_n_cand, S.sc, S.mx.p, S.fail.p,
              ctx->world > 1 ? nullptr : ctx->h_scalars_dev, ctx->world > 1 ? 0.0 : ++ctx->h_seq,
              ctx->world > 1 ? S.gat.p : nullptr, ctx->rank, ctx->world};
  LAUNCH(launch_k(iteration_tail_kernel, 1, 256, 0, st, ta));
  TSL_CHECK_LAUNCH();
  mark(S, 7);  // end of the iteration's device work
  return TSLAM_OK;
}


// Spins on the sequence number the device writes after the iteration's scalars (page-locked, device-mapped host memory);
// the stream is queried now and then so that a failed launch or a sticky error ends the wait instead of hanging it.
static int wait_published(tslam_ctx* ctx) {
  volatile double* flag = ctx->h_scalars + H_SEQ;
  const double seq = ctx->h_seq;
  for (unsigned spins = 1;; ++spins) {
    if (*flag == seq) break;
    if ((spins & 0x1fffu) == 0) {
      const cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e == cudaErrorNotReady) continue;
      if (e != cudaSuccess) return set_error(TSLAM_ERR_CUDA, "LM iteration failed on the device: %s", cudaGetErrorString(e));
      if (*flag == seq) break;
      return set_error(TSLAM_ERR_CUDA, "LM iteration finished without publishing its scalars");
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return TSLAM_OK;
}

static int run_lm(Solver& S, const tslam_solve_options* opt, int max_iters, tslam_solve_summary* out, double* trace) {
  tslam_ctx* ctx = S.ctx; tslam_dev_problem* d = S.d; cudaStream_t st = ctx->stream;
  auto T0 = std::chrono::steady_clock::now();
  const double ftol = opt->function_tolerance > 0 ? opt->function_tolerance : 1e-6;
  const double gtol = opt->gradient_tolerance > 0 ? opt->gradient_tolerance : 1e-10;
  const double ptol = opt->parameter_tolerance > 0 ? opt->parameter_tolerance : 1e-8;
  double radius = opt->initial_radius > 0 ? opt->initial_radius : 1e4;
  const double max_radius = 1e16, min_radius = 1e-32, min_rel_dec = 1e-3;
  double decrease_factor = 2.0;
  const int jac_mode = opt->text_jac_mode;
  tslam_solve_summary sum{};
  sum.n_free_cams = S.nc; sum.n_free_points = S.nl; sum.n_free_planes = S.npl; sum.reduced_dim = S.n; sum.setup_ms = S.setup_ms;
  double* h = ctx->h_scalars;

  // current / candidate parameter buffers
  S.x_cams = d->cams.p; S.x_rho = d->rho.p; S.x_theta = d->theta.p;
  S.c_cams = S.xc_cams.p; S.c_rho = S.xc_rho.p; S.c_theta = S.xc_theta.p;
  TSL_CUDA(cudaMemcpyAsync(S.c_rho, S.x_rho, sizeof(double) * d->n_points, cudaMemcpyDeviceToDevice, st));
  TSL_CUDA(cudaMemcpyAsync(S.c_theta, S.x_theta, sizeof(double) * 3 * (size_t)d->n_planes, cudaMemcpyDeviceToDevice, st));

  TSL_CUDA(cudaMemsetAsync(S.mx.p, 0, MX_N * sizeof(double), st));
  TSL_CUDA(cudaMemsetAsync(S.fail.p, 0, sizeof(int), st));
  S.rp = S.pr.p; S.Jp = S.pJ.p; S.rt = S.tr.p; S.Jt = S.tJ.p; S.rp2 = S.cr_p.p; S.Jp2 = S.pJ2.p; S.rt2 = S.cr_t.p; S.Jt2 = S.tJ2.p;
  // Speculative Jacobian: worth it when a Jacobian costs about as much as a residual pass (closed forms); with Ceres-style
  // central differences (35 functor calls per text block) a rejected step would waste far more than an accepted one saves.
  {
    static const bool spec_env = [] { const char* e = getenv("TSLAM_SPEC_J"); return !(e && e[0] == '0'); }();
    S.spec_J = spec_env && (S.lt == 0 || jac_mode == TSLAM_JAC_ANALYTIC);
  }
  // ---- iteration 0 ----
  mark(S, 0);
  int rc = eval_at(S, S.x_cams, S.x_rho, S.x_theta, true, jac_mode, SC_COST, S.rp, S.rt, S.Jp, S.Jt);
  if (rc) return rc;
  if ((rc = compute_jacobi_scaling(S))) return rc;
  if ((rc = accumulate_landmarks(S))) return rc;
  // x_norm
  {
    LAUNCH(launch_k(xnorm_cams_kernel, 1, 256, 0, st, S.K, S.camslot_d.p, S.x_cams, S.sc + SC_XNORM2, ctx->rank == 0 ? 1.0 : 0.0));
    const int gvp = grid_for(S.nvp, 256), gvt = grid_for(S.nvt, 256);
    if (S.nvp) LAUNCH(launch_k(xnorm_lm_kernel<1>, gvp, 256, 0, st, S.nvp, S.vp_gl.p, S.x_rho, S.parts.p));
    if (S.nvt) LAUNCH(launch_k(xnorm_lm_kernel<3>, gvt, 256, 0, st, S.nvt, S.vt_gl.p, S.x_theta, S.parts.p + gvp));
    if (gvp + gvt) LAUNCH(launch_k(sum_parts_kernel, 1, 256, 0, st, S.parts.p, gvp + gvt, 1, 0, S.sc + SC_XNORM2, 1));
    TSL_CHECK_LAUNCH();
  }
  double x_cost = 0, fixed_cost = 0, x_norm = 0, gmax = 0;
  int iter = 0, n_ok = 0, n_bad = 0, invalid_run = 0, term = TSLAM_TERM_NO_CONVERGENCE;
  bool have_cost = false;
  auto T1 = std::chrono::steady_clock::now();

  while (true) {
    if (iter >= max_iters) { term = TSLAM_TERM_NO_CONVERGENCE; break; }
    if (radius <= min_radius) { term = TSLAM_TERM_NO_CONVERGENCE; break; }
    if (S.nc + S.nl + S.npl == 0) { term = TSLAM_TERM_GRADIENT_TOL; break; }   // nothing to optimise: the (empty) gradient passes Ceres' first test
    // ---- linear solve + candidate (speculative: the gradient test for THIS iteration is read back with it) ----
    if ((rc = compute_step(S, radius))) return rc;
    // (the gradient max-norm pieces run inside post_solve_kernel)
    if ((rc = model_and_candidate_cost(S, jac_mode))) return rc;
    if (ctx->world > 1) {  // ONE small collective per iteration for the 8 sums and 2 maxima (iteration_tail_kernel filled this rank's slot)
      if ((rc = comm_allreduce_sum(ctx, S.gat.p, (size_t)ctx->world * (SC_N + MX_N)))) return rc;
      LAUNCH(launch_k(publish_gathered_kernel, 1, 32, 0, st, S.gat.p, ctx->world, S.mx.p, S.fail.p, ctx->h_scalars_dev, ++ctx->h_seq));
      TSL_CHECK_LAUNCH();
    }
    if ((rc = wait_published(ctx))) return rc;   // iteration_tail_kernel / publish_scalars_kernel wrote h[] and the sequence number
    if (!have_cost) {
      x_cost = h[SC_COST]; fixed_cost = h[SC_FIXED]; x_norm = std::sqrt(h[SC_XNORM2]); have_cost = true;
      sum.initial_cost = x_cost + fixed_cost; sum.fixed_cost = fixed_cost;
      if (trace) { trace[0] = x_cost + fixed_cost; trace[1] = radius; trace[2] = 0; trace[3] = 1; }
    }
    gmax = h[SC_N + MX_GMAX];
    if (gmax <= gtol) { term = TSLAM_TERM_GRADIENT_TOL; break; }  // Ceres tests this before computing the step
    ++iter;
    const bool lin_fail = h[SC_N + MX_FAIL] != 0.0;
    const double mcc = h[SC_MCC], cand_cost = h[SC_CAND], step_norm = std::sqrt(h[SC_STEP2]);
    const bool finite = std::isfinite(mcc) && std::isfinite(cand_cost) && std::isfinite(step_norm);
    if (lin_fail || !finite || !(mcc > 0.0)) {
      ++invalid_run; ++n_bad;
      if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = 0; t[3] = -1; }
      if (invalid_run >= 5) { term = TSLAM_TERM_FAILURE; break; }
      radius /= decrease_factor; decrease_factor *= 2.0;
      continue;
    }
    invalid_run = 0;
    if (step_norm <= ptol * (x_norm + ptol)) {
      term = TSLAM_TERM_PARAMETER_TOL;
      if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = 0; t[3] = 0; }
      break;
    }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= ftol * x_cost) {
      term = TSLAM_TERM_FUNCTION_TOL;
      if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = 0; t[3] = 0; }
      break;
    }
    const double rel = cost_change / mcc;
    if (rel > min_rel_dec) {
      // both buffers carry the untouched (fixed / non-owned) landmark entries since the copy before iteration 0, and
      // post_solve_kernel rewrites every owned free landmark entry each iteration: a pointer swap is all an accepted step needs
      std::swap(S.x_cams, S.c_cams); std::swap(S.x_rho, S.c_rho); std::swap(S.x_theta, S.c_theta);
      x_norm = std::sqrt(h[SC_CNORM2]);
      mark(S, 0);  // eval + J (the next iteration's linearisation point)
      if (S.spec_J) {   // residuals and Jacobian of the accepted point were produced by the candidate evaluation
        std::swap(S.rp, S.rp2); std::swap(S.Jp, S.Jp2); std::swap(S.rt, S.rt2); std::swap(S.Jt, S.Jt2);
      } else if ((rc = eval_at(S, S.x_cams, S.x_rho, S.x_theta, true, jac_mode, SC_COST, S.rp, S.rt, S.Jp, S.Jt, /*defer_sum=*/true))) return rc;
      if ((rc = accumulate_landmarks(S))) return rc;
      x_cost = cand_cost;  // identical evaluation point; the device value is re-read next iteration for the trace only
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(max_radius, radius); decrease_factor = 2.0;
      ++n_ok;
      if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = rel; t[3] = 1; }
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0;
      ++n_bad;
      if (trace) { double* t = trace + 4 * iter; t[0] = cand_cost + fixed_cost; t[1] = radius; t[2] = rel; t[3] = 0; }
    }
  }
  TSL_CUDA(cudaStreamSynchronize(st));
  if (!have_cost) {   // left before the first read-back (no free parameter block, or max_iters == 0): the costs of iteration 0 are still wanted
    const double* sc_src = S.sc;
    if (ctx->world > 1) {
      TSL_CUDA(cudaMemcpyAsync(S.scr.p, S.sc, SC_N * sizeof(double), cudaMemcpyDeviceToDevice, st));
      if ((rc = comm_allreduce_sum(ctx, S.scr.p, SC_N))) return rc;
      sc_src = S.scr.p;
    }
    TSL_CUDA(cudaMemcpyAsync(h, sc_src, SC_N * sizeof(double), cudaMemcpyDeviceToHost, st));
    TSL_CUDA(cudaStreamSynchronize(st));
    x_cost = h[SC_COST]; fixed_cost = h[SC_FIXED];
    sum.initial_cost = x_cost + fixed_cost; sum.fixed_cost = fixed_cost;
    if (trace) { trace[0] = x_cost + fixed_cost; trace[1] = radius; trace[2] = 0; trace[3] = 1; }
  }
  // make d->cams / rho / theta hold the solution
  if (S.x_cams != d->cams.p) {
    TSL_CUDA(cudaMemcpyAsync(d->cams.p, S.x_cams, sizeof(double) * 7 * (size_t)S.K, cudaMemcpyDeviceToDevice, st));
    TSL_CUDA(cudaMemcpyAsync(d->rho.p, S.x_rho, sizeof(double) * d->n_points, cudaMemcpyDeviceToDevice, st));
    TSL_CUDA(cudaMemcpyAsync(d->theta.p, S.x_theta, sizeof(double) * 3 * (size_t)d->n_planes, cudaMemcpyDeviceToDevice, st));
    TSL_CUDA(cudaStreamSynchronize(st));
    S.x_cams = d->cams.p; S.x_rho = d->rho.p; S.x_theta = d->theta.p;
  }
  auto T2 = std::chrono::steady_clock::now();
  sum.iterations = iter; sum.successful_steps = n_ok; sum.unsuccessful_steps = n_bad; sum.termination = term;
  sum.final_cost = x_cost + fixed_cost;
  sum.solve_ms = std::chrono::duration<double, std::milli>(T2 - T1).count() + std::chrono::duration<double, std::milli>(T1 - T0).count();
  if (out) *out = sum;
  return TSLAM_OK;
}

void free_solver(tslam_dev_problem* d) {
  if (d && d->solver) { delete static_cast<Solver*>(d->solver); d->solver = nullptr; }
}

static int get_solver(tslam_ctx* ctx, tslam_dev_problem* d, Solver** out) {
  if (!d->solver) {
    Solver* S = new Solver();
    S->ctx = ctx; S->d = d;
    int rc = analyze_and_upload(*S);
    if (rc) { delete S; return rc; }
    d->solver = S;
  }
  *out = static_cast<Solver*>(d->solver);
  return TSLAM_OK;
}

}  // namespace tsl

using namespace tsl;

namespace tsl {
int gate_device(tslam_ctx* ctx, const double* d_rp, const double* d_rt, int n_pobs, int n_tobs, const int32_t* t_obj,
                const int32_t* obj_size, int n_obj, const tslam_gate_options* g, uint8_t* pt_bad, uint8_t* tf_bad, uint8_t* obj_bad,
                int32_t* counts_out);   // gate.cu
struct GateArgs {
  const tslam_gate_options* opt; const int32_t* t_obj; const int32_t* obj_size; int n_obj;
  uint8_t *pt_bad, *tf_bad, *obj_bad; int32_t* counts;
};
}  // namespace tsl

static int solve_impl(tslam_ctx* ctx, tslam_ba_problem* p, const tslam_solve_options* opt, tslam_solve_summary* summary,
                      double* final_residuals, double* trace, const GateArgs* gate) {
  if (!ctx || !p || !opt) return set_error(TSLAM_ERR_ARG, "null argument");
  if (opt->max_iters < 0) return set_error(TSLAM_ERR_ARG, "max_iters < 0");
  auto T0 = std::chrono::steady_clock::now();
  TSL_CUDA(cudaSetDevice(ctx->device));
  AllocStreamScope alloc_scope(ctx->stream);   // declared before every device buffer of this call: released after them, on the same stream
  int rc = validate_problem(p);
  if (rc) return rc;
  {   // pose-only / local-window sizes: one persistent kernel for the whole solve (ba_small.cu)
    bool handled = false;
    const double *d_rp = nullptr, *d_rt = nullptr;
    tslam_solve_summary ssum{};
    if ((rc = small_solve(ctx, p, opt, &ssum, final_residuals, trace, &d_rp, &d_rt, &handled))) return rc;
    if (handled) {
      if (gate && (rc = gate_device(ctx, d_rp, d_rt, p->n_pobs, p->n_tobs, gate->t_obj, gate->obj_size, gate->n_obj, gate->opt, gate->pt_bad, gate->tf_bad,
                                    gate->obj_bad, gate->counts))) return rc;
      ssum.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
      if (summary) *summary = ssum;
      return TSLAM_OK;
    }
  }
  tslam_dev_problem d;
  rc = upload_problem(ctx, p, &d, /*shard=*/true, /*persistent=*/false, /*validated=*/true);
  if (rc) return rc;
  auto Tu = std::chrono::steady_clock::now();
  struct Guard { tslam_dev_problem* d; ~Guard() { free_solver(d); } } guard{&d};
  Solver* S = nullptr;
  if ((rc = get_solver(ctx, &d, &S))) return rc;
  auto T1 = std::chrono::steady_clock::now();
  tslam_solve_summary sum{};
  if ((rc = run_lm(*S, opt, opt->max_iters, &sum, trace))) return rc;
  auto Tlm = std::chrono::steady_clock::now();
  cudaStream_t st = ctx->stream;
  // ---- results back to the caller's arrays ----
  if (ctx->world > 1) {
    // every rank owns a subset of the landmarks: one packed buffer [rho | theta | owner count per point | per plane], zero where
    // not owned, summed across the ranks; landmarks nobody owns (constant / unobserved) keep the caller's values
    const size_t nP = (size_t)d.n_points, nT = (size_t)d.n_planes, total = 2 * nP + 4 * nT;
    DevBuf<double> ob;
    TSL_CUDA(ob.reserve(total));
    TSL_CUDA(cudaMemsetAsync(ob.p, 0, sizeof(double) * (total ? total : 1), st));
    double *orho = ob.p, *oth = ob.p + nP, *ownP = ob.p + nP + 3 * nT, *ownT = ownP + nP;
    if (S->nvp) LAUNCH(launch_k(export_lm_kernel<1>, (S->nvp + 255) / 256, 256, 0, st, S->nvp, S->vp_gl.p, d.rho.p, orho, ownP));
    if (S->nvt) LAUNCH(launch_k(export_lm_kernel<3>, (3 * S->nvt + 255) / 256, 256, 0, st, S->nvt, S->vt_gl.p, d.theta.p, oth, ownT));
    if ((rc = comm_allreduce_sum(ctx, ob.p, total))) return rc;
    if (nP) LAUNCH(launch_k(merge_lm_kernel<1>, (int)((nP + 255) / 256), 256, 0, st, (int)nP, orho, ownP, d.rho.p));
    if (nT) LAUNCH(launch_k(merge_lm_kernel<3>, (int)((3 * nT + 255) / 256), 256, 0, st, (int)nT, oth, ownT, d.theta.p));
    TSL_CHECK_LAUNCH();
    TSL_CUDA(cudaMemcpyAsync(p->cams, d.cams.p, sizeof(double) * 7 * (size_t)d.n_cams, cudaMemcpyDeviceToHost, st));
    if (nP) TSL_CUDA(cudaMemcpyAsync(p->rho, d.rho.p, sizeof(double) * nP, cudaMemcpyDeviceToHost, st));
    if (nT) TSL_CUDA(cudaMemcpyAsync(p->theta, d.theta.p, sizeof(double) * 3 * nT, cudaMemcpyDeviceToHost, st));
    TSL_CUDA(cudaStreamSynchronize(st));
  } else {
    TSL_CUDA(cudaMemcpyAsync(p->cams, d.cams.p, sizeof(double) * 7 * (size_t)d.n_cams, cudaMemcpyDeviceToHost, st));
    if (d.n_points) TSL_CUDA(cudaMemcpyAsync(p->rho, d.rho.p, sizeof(double) * d.n_points, cudaMemcpyDeviceToHost, st));
    if (d.n_planes) TSL_CUDA(cudaMemcpyAsync(p->theta, d.theta.p, sizeof(double) * 3 * (size_t)d.n_planes, cudaMemcpyDeviceToHost, st));
  }
  DevBuf<double> fr;   // world > 1: the global residual vector, replicated after the all-reduce
  if (final_residuals || gate) {
    // Problem::Evaluate: loss-corrected residuals of every block, insertion order (Appendix A.6)
    if ((rc = eval_at(*S, d.cams.p, d.rho.p, d.theta.p, false, opt->text_jac_mode, SC_CAND, S->rp2, S->rt2, nullptr, nullptr))) return rc;
    const size_t total = 2 * (size_t)d.g_pobs + 8 * (size_t)d.g_tobs;
    if (ctx->world > 1) {
      TSL_CUDA(fr.reserve(total));
      TSL_CUDA(cudaMemsetAsync(fr.p, 0, total * sizeof(double), st));
      if (S->lp) LAUNCH(launch_k(scatter_rows_kernel, (2 * S->lp + 255) / 256, 256, 0, st, S->lp, 2, S->gsel_p.p, S->rp2, fr.p));
      if (S->lt) LAUNCH(launch_k(scatter_rows_kernel, (8 * S->lt + 255) / 256, 256, 0, st, S->lt, 8, S->gsel_t.p, S->rt2, fr.p + 2 * (size_t)d.g_pobs));
      if ((rc = comm_allreduce_sum(ctx, fr.p, total))) return rc;
      if (final_residuals) TSL_CUDA(cudaMemcpyAsync(final_residuals, fr.p, total * sizeof(double), cudaMemcpyDeviceToHost, st));
      TSL_CUDA(cudaStreamSynchronize(st));
    } else if (final_residuals) {
      if (S->lp) TSL_CUDA(cudaMemcpyAsync(final_residuals, S->rp2, 2 * (size_t)S->lp * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (S->lt) TSL_CUDA(cudaMemcpyAsync(final_residuals + 2 * (size_t)d.g_pobs, S->rt2, 8 * (size_t)S->lt * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (gate) {   // chi^2 gates on the residuals where they are (src/optimizer.cc:1236-1302, 1616-1684)
      const double* rp = ctx->world > 1 ? fr.p : S->rp2;
      const double* rt = ctx->world > 1 ? fr.p + 2 * (size_t)d.g_pobs : S->rt2;
      if ((rc = gate_device(ctx, rp, rt, d.g_pobs, d.g_tobs, gate->t_obj, gate->obj_size, gate->n_obj, gate->opt, gate->pt_bad, gate->tf_bad,
                            gate->obj_bad, gate->counts))) return rc;
    }
  }
  TSL_CUDA(cudaStreamSynchronize(st));
  auto T2 = std::chrono::steady_clock::now();
  if (getenv("TSLAM_SETUP_TRACE"))
    fprintf(stderr, "[tslam solve] upload %.3f ms, analysis %.3f ms, lm %.3f ms, download %.3f ms\n", std::chrono::duration<double, std::milli>(Tu - T0).count(),
            std::chrono::duration<double, std::milli>(T1 - Tu).count(), std::chrono::duration<double, std::milli>(Tlm - T1).count(),
            std::chrono::duration<double, std::milli>(T2 - Tlm).count());
  sum.setup_ms = std::chrono::duration<double, std::milli>(T1 - T0).count();
  sum.total_ms = std::chrono::duration<double, std::milli>(T2 - T0).count();
  if (summary) *summary = sum;
  return TSLAM_OK;
}

extern "C" int tslam_solve(tslam_ctx* ctx, tslam_ba_problem* p, const tslam_solve_options* opt, tslam_solve_summary* summary,
                           double* final_residuals, double* trace) {
  return solve_impl(ctx, p, opt, summary, final_residuals, trace, nullptr);
}

extern "C" int tslam_solve_gated(tslam_ctx* ctx, tslam_ba_problem* p, const tslam_solve_options* opt, const tslam_gate_options* gate,
                                 const int32_t* t_obj, const int32_t* obj_size, int n_obj, tslam_solve_summary* summary,
                                 double* final_residuals, double* trace, uint8_t* pt_bad, uint8_t* tf_bad, uint8_t* obj_bad,
                                 int32_t counts_out[3]) {
  if (!gate) return set_error(TSLAM_ERR_ARG, "gate options missing");
  if (counts_out) counts_out[0] = counts_out[1] = counts_out[2] = 0;
  GateArgs g{gate, t_obj, obj_size, n_obj, pt_bad, tf_bad, obj_bad, counts_out};
  return solve_impl(ctx, p, opt, summary, final_residuals, trace, &g);
}

extern "C" int tslam_dev_lm_iterations(tslam_ctx* ctx, tslam_dev_problem* d, const tslam_solve_options* opt, int iters, float* phase_ms,
                                       tslam_solve_summary* summary) {
  if (!ctx || !d || !opt) return set_error(TSLAM_ERR_ARG, "null argument");
  TSL_CUDA(cudaSetDevice(ctx->device));
  Solver* S = nullptr;
  int rc = get_solver(ctx, d, &S);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  // reset parameters to the uploaded values
  TSL_CUDA(cudaMemcpyAsync(d->cams.p, d->cams0.p, sizeof(double) * 7 * (size_t)d->n_cams, cudaMemcpyDeviceToDevice, st));
  TSL_CUDA(cudaMemcpyAsync(d->rho.p, d->rho0.p, sizeof(double) * d->n_points, cudaMemcpyDeviceToDevice, st));
  TSL_CUDA(cudaMemcpyAsync(d->theta.p, d->theta0.p, sizeof(double) * 3 * (size_t)d->n_planes, cudaMemcpyDeviceToDevice, st));
  S->timing = phase_ms != nullptr; S->ev_used = 0;
  tslam_solve_summary sum{};
  TSL_CUDA(cudaEventRecord(ctx->ev0, st));
  rc = run_lm(*S, opt, iters, &sum, nullptr);
  if (rc) return rc;
  TSL_CUDA(cudaEventRecord(ctx->ev1, st));
  TSL_CUDA(cudaEventSynchronize(ctx->ev1));
  if (phase_ms) {
    for (int k = 0; k < 8; ++k) phase_ms[k] = 0.f;
    float whole = 0.f;
    cudaEventElapsedTime(&whole, ctx->ev0, ctx->ev1);
    for (int k = 0; k + 1 < S->ev_used; ++k) {
      const int ph = S->ev_kind[k];
      if (ph < 0 || ph > 6) continue;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, S->ev[k], S->ev[k + 1]);
      phase_ms[ph] += ms;
    }
    const int denom = sum.iterations > 0 ? sum.iterations : 1;
    for (int k = 0; k < 7; ++k) phase_ms[k] /= denom;
    phase_ms[7] = whole / denom;
  }
  S->timing = false;
  if (summary) *summary = sum;
  return TSLAM_OK;
}
