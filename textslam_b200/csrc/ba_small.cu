// Small-problem solve: ONE persistent cooperative kernel per tslam_solve call (pose-only tracking, local BA windows).
//
// The reference solves these inside ceres::Solve at src/optimizer.cc:1218 (PyrPoseOptim: one free pose, ~2k point + ~250 text
// blocks) and :1600 (PyrBA: <= 10 free poses); it does so three times per frame / key frame (levels 2, 1, 0,
// src/optimizer.cc:172-186, 282-289), so the fixed cost of a call is what matters, not throughput. The general path of ba_solve.cu
// pays ~20 uploads, a structure analysis with a device round trip, ~130 stream-ordered allocations and ~14 launches plus one
// host hand-shake per LM iteration; on these sizes that is all latency. Here:
//   host   one O(n) pass over the index arrays (free-camera slots, observation lists per free landmark), everything packed into
//          ONE page-locked staging block -> one H2D copy (+ one for the images), one kernel, one D2H copy of the result block;
//   device the whole Levenberg-Marquardt loop of run_lm() (same accept / reject / termination rules, same Jacobi-scaled damping)
//          in one cooperative launch of G CTAs (one per SM as soon as there is work for it) that meet at grid barriers: evaluate r, J -> per-warp accumulation of the camera block
//          H_cc and of every landmark's V, g, E (observation lists, fixed order) -> Schur complement contributions -> fixed-order
//          reduction over CTAs -> dense Cholesky of the <= 60 x 60 reduced camera system in shared memory (every CTA redundantly:
//          no broadcast) -> back-substitution, candidate, model cost change, candidate evaluation (with its Jacobian,
//          speculatively) -> decision, taken identically by every CTA from the same partial sums.
// All sums are taken in a fixed order: results are reproducible run to run.
// Algebra: the Jacobi scaling s of Ceres enters only through the damping term; with M_l = V_l + diag(clamp(s^2 V_ii)/(radius s^2))
//   (H_cc + D_c - sum_l E_l^T M_l^-1 E_l) u = g_c - sum_l E_l^T M_l^-1 g_l,  delta_c = -u,  delta_l = -M_l^-1 (g_l + E_l delta_c)
// is the scaled system of ba_solve.cu / oracle/ba_lm.cpp written in unscaled variables.
// Eligible: one GPU, analytic text Jacobian, <= 10 free cameras, <= 256 cameras, <= 16384 point and <= 4096 text blocks, no
// landmark with more than 1024 observations. Anything else takes the general path (TSLAM_SMALL=0 forces that).
#include "ctx.cuh"
#include "ba_device.cuh"
#include "solver.cuh"
#include "chol_common.cuh"
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace tsl {

constexpr int ST = 256, SW = 8;
constexpr int S_MAX_NC = 10, S_MAX_K = 256, S_MAX_P = 16384, S_MAX_T = 4096, S_MAX_LIST = 1024, S_MAX_G = 160;
constexpr int S_EC = 64;            // columns of a landmark's E row (6 * S_MAX_NC padded)
constexpr int S_SCR = 848;          // doubles of scratch per warp: [0, 448) staged observation blocks, [448, 832) E / W rows, [832, 848) slot list
constexpr int S_STG = 448, S_CHUNK = 16;
constexpr int S_SPIN = 1 << 22;
constexpr int S_NPART = 8;

struct SmallArgs {
  // ---- launch-invariant inputs ----
  const double2 *p_uv, *p_ray;
  const int *p_cam, *p_host, *p_lm, *p_ls, *p_cs, *p_hs;   // *_cs / *_hs: free-camera slot of the observing / host camera (-1: constant)
  const double2 *t_rays, *t_musigma;
  const double* t_iref;
  const int *t_cam, *t_host, *t_plane, *t_img, *t_ls, *t_cs, *t_hs;
  const uint8_t* imgs; int img_w, img_h;
  const int* camslot;
  const int *vp_ptr, *vp_obs, *vp_gl, *vt_ptr, *vt_obs, *vt_gl;
  double pfx, pfy, pcx, pcy, wx, wy, hub_p, tfx, tfy, tcx, tcy, wT, hub_t;
  // ---- mutable state ----
  const double* cams_in;
  double *rho[2], *theta[2];
  double *rp[2], *Jp[2], *rt[2], *Jt[2];
  double *Vp, *gp, *Ep, *Mp, *sclp, *dlp;       // per free inverse depth: V, g, E (64), M^-1, Jacobi scale, step
  double *Vt, *gt, *Et, *Mt, *sclt, *dlt;       // per free plane: V (6: 00 01 02 11 12 22), g (3), E (3 x 64), M^-1 (6), scale (3), step (3)
  unsigned *maskp, *maskt;                      // camera slots a landmark touches
  double *partH, *partS, *sumH, *sumS, *part0, *partG, *part;
  int* sync;                                    // [0] barrier counter, [1] abort
  double* out;                                  // result block, see OUT_*
  long long* prof;                              // TSLAM_SMALL_PROF: clock64 stamps of CTA 0, 16 per iteration
  double *r_final_p, *r_final_t;
  // ---- sizes / options ----
  int K, nc, n, lp, lt, nvp, nvt, n_points, n_planes, G, max_iters;
  double ftol, gtol, ptol, radius0;
};
enum { OUT_ITER = 0, OUT_OK, OUT_BAD, OUT_TERM, OUT_INIT, OUT_FINAL, OUT_FIXED, OUT_ABORT, OUT_HDR = 8 };

__device__ __forceinline__ int ld_acq(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier on a monotonic counter (zeroed by the upload that precedes the launch). Returns true when the solve was aborted
// (a CTA waited too long: a bug, never a data-dependent condition) — every CTA then leaves the kernel.
struct GridBar { int* ctr; int* abort; int G; int epoch; };
__device__ __forceinline__ bool grid_sync(GridBar& b) {
  __shared__ int s_abort;
  __syncthreads();
  if (threadIdx.x == 0) {
    b.epoch += b.G;
    __threadfence();
    atomicAdd(b.ctr, 1);
    int spins = 0, ab = 0;
    while (ld_acq(b.ctr) < b.epoch) {
      if (++spins > S_SPIN) { ab = 1; atomicExch(b.abort, 1); break; }
      if ((spins & 255) == 0 && ld_acq(b.abort)) { ab = 1; break; }
    }
    __threadfence();
    s_abort = ab;
  }
  __syncthreads();
  return s_abort != 0;
}

__device__ __forceinline__ Cam cam_at(const double* cams, int k) {
  Cam c;
  const double* p = cams + 7 * k;
#pragma unroll
  for (int i = 0; i < 4; ++i) c.q[i] = p[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) c.t[i] = p[4 + i];
  return c;
}

// deterministic CTA sums of NV values (shuffle tree per warp, then warps in order); result valid on thread 0
template <int NV>
__device__ __forceinline__ void block_sums(double (&v)[NV], double* s_red) {
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) s_red[(threadIdx.x >> 5) * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) { double s = 0.0; for (int w = 0; w < SW; ++w) s += s_red[w * NV + k]; v[k] = s; }
}

// sum / max over the G per-CTA partials base[c * stride], c = lane, lane + 32, ... then a shuffle tree: the loads are independent (one
// or two L2 round trips instead of G dependent ones) and the order of the additions is fixed. Result in every lane.
__device__ __forceinline__ double warp_sum_ctas(const double* base, int stride, int G, int lane) {
  double s = 0.0;
  for (int c = lane; c < G; c += 32) s += base[(size_t)c * stride];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
__device__ __forceinline__ double warp_max_ctas(const double* base, int stride, int G, int lane) {
  double s = 0.0;
  for (int c = lane; c < G; c += 32) s = fmax(s, base[(size_t)c * stride]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}
#define STAMP(k) do { if (a.prof && blockIdx.x == 0 && threadIdx.x == 0 && pit < 16) a.prof[16 * pit + (k)] = clock64(); } while (0)

// residuals (+ Jacobians) of every block at (cams in shared memory, rho, theta): loss-corrected, observation-major
template <bool WJ>
__device__ __forceinline__ void eval_pass(const SmallArgs& a, const double* scams, const double* rho, const double* theta, double* rp, double* Jp,
                                          double* rt, double* Jt, double& cost_a, double& cost_f) {
  // one index space: [0, lp) point blocks, then (from the next multiple of 32) the 8 lt text pixels; a warp is all points or all
  // text, and with G x 256 threads >= lp + 8 lt a thread evaluates one item (the two latency chains run side by side)
  const int lp_pad = (a.lp + 31) & ~31;
  for (int idx = blockIdx.x * ST + threadIdx.x; idx < lp_pad + ((8 * a.lt + 31) & ~31); idx += a.G * ST) {
    if (idx < lp_pad) {
      const int i = idx;
      if (i >= a.lp) continue;
      const double2 uv = a.p_uv[i], ray = a.p_ray[i];
      const int ci = a.p_cam[i], hi = a.p_host[i], li = a.p_lm[i];
      const Cam c = cam_at(scams, ci), h = cam_at(scams, hi);
      double r[2], J[26];
      point_eval<WJ>(c, h, rho[li], ray.x, ray.y, uv.x, uv.y, a.pfx, a.pfy, a.pcx, a.pcy, a.wx, a.wy, r, J);
      double rho0;
      const double sq = huber_scale(a.hub_p, r[0] * r[0] + r[1] * r[1], &rho0);
      const bool act = a.p_cs[i] >= 0 || a.p_hs[i] >= 0 || a.p_ls[i] >= 0;
      if (act) cost_a += 0.5 * rho0; else cost_f += 0.5 * rho0;
      rp[2 * i] = r[0] * sq; rp[2 * i + 1] = r[1] * sq;
      if (WJ) {
        double* o = Jp + 26 * (size_t)i;
#pragma unroll
        for (int k = 0; k < 26; ++k) o[k] = J[k] * sq;
      }
      continue;
    }
    const int gpx = idx - lp_pad;
    const int b = gpx >> 3;
    const bool valid = b < a.lt;
    double res = 0.0, Jr[15];
    bool act = false;
    if (valid) {
      const int ci = a.t_cam[b], hi = a.t_host[b], pl = a.t_plane[b];
      const Cam c = cam_at(scams, ci), h = cam_at(scams, hi);
      const double th[3] = {theta[3 * pl], theta[3 * pl + 1], theta[3 * pl + 2]};
      const double2 ray = a.t_rays[gpx];
      const double2 ms = a.t_musigma[b];
      TextImg im{a.imgs + (size_t)a.t_img[b] * a.img_w * a.img_h, a.img_w, a.img_h};
      if (WJ) res = text_pixel_analytic(c, h, th, ray.x, ray.y, im, a.tfx, a.tfy, a.tcx, a.tcy, ms.x, ms.y, a.t_iref[gpx], a.wT, Jr);
      else res = text_residual_only(c.q, c.t, h.q, h.t, th, ray.x, ray.y, im, a.tfx, a.tfy, a.tcx, a.tcy, ms.x, ms.y, a.t_iref[gpx], a.wT);
      act = a.t_cs[b] >= 0 || a.t_hs[b] >= 0 || a.t_ls[b] >= 0;
    }
    double s = res * res;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    double rho0;
    const double sq = huber_scale(a.hub_t, s, &rho0);
    if (valid) {
      if ((gpx & 7) == 0) { if (act) cost_a += 0.5 * rho0; else cost_f += 0.5 * rho0; }
      rt[gpx] = res * sq;
      if (WJ) {
        double* o = Jt + 15 * (size_t)gpx;
#pragma unroll
        for (int k = 0; k < 15; ++k) o[k] = Jr[k] * sq;
      }
    }
  }
}

// One residual block (R rows of JC Jacobian columns, then the R residuals) travels global -> registers -> this warp's scratch;
// the loads of the next block are issued before the current one is consumed.
template <int R, int JC>
struct BlockRegs {
  static constexpr int NV = (R * JC + R + 31) / 32;
  double v[NV];
  __device__ __forceinline__ void load(const double* J, const double* r, int lane) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int q = lane + 32 * j;
      v[j] = q < R * JC ? J[q] : (q < R * JC + R ? r[q - R * JC] : 0.0);
    }
  }
  __device__ __forceinline__ void store(double* scr, int lane) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) { const int q = lane + 32 * j; if (q < R * JC + R) scr[q] = v[j]; }
  }
};

// camera part of the staged block into this warp's packed accumulator: lower triangle of H_cc, then g_c
template <int R, int JC>
__device__ __forceinline__ void accum_block(double* acc, const double* scr, int cs, int hs, int NT, int lane) {
  for (int e = lane; e < 144; e += 32) {
    const int ia = e / 12, ib = e - 12 * ia;
    const int sa = ia < 6 ? cs : hs, sb = ib < 6 ? cs : hs;
    if (sa < 0 || sb < 0) continue;
    const int I = 6 * sa + (ia < 6 ? ia : ia - 6), Jx = 6 * sb + (ib < 6 ? ib : ib - 6);
    if (I < Jx) continue;
    double v = 0.0;
#pragma unroll
    for (int row = 0; row < R; ++row) v += scr[row * JC + ia] * scr[row * JC + ib];
    acc[I * (I + 1) / 2 + Jx] += v;
  }
  if (lane < 12) {
    const int s = lane < 6 ? cs : hs;
    if (s >= 0) {
      double v = 0.0;
#pragma unroll
      for (int row = 0; row < R; ++row) v += scr[row * JC + lane] * scr[R * JC + row];
      acc[NT + 6 * s + (lane < 6 ? lane : lane - 6)] += v;
    }
  }
}

template <int R, int JC>
__device__ __forceinline__ void accum_pass(double* acc, double* scr, int first, int count, int stride, const int* cs_of, const int* hs_of, const double* J,
                                           const double* r, int NT, int lane) {
  BlockRegs<R, JC> regs;
  int i = first;
  int cs = -1, hs = -1;
  if (i < count) { cs = cs_of[i]; hs = hs_of[i]; regs.load(J + (size_t)R * JC * i, r + R * i, lane); }
  while (i < count) {
    const int ccs = cs, chs = hs;
    regs.store(scr, lane);
    __syncwarp();
    const int nx = i + stride;
    if (nx < count) { cs = cs_of[nx]; hs = hs_of[nx]; regs.load(J + (size_t)R * JC * nx, r + R * nx, lane); }
    if (ccs >= 0 || chs >= 0) accum_block<R, JC>(acc, scr, ccs, chs, NT, lane);
    __syncwarp();
    i = nx;
  }
}

// 1 / d for d > 0 without the division subroutine: hardware seed (rcp.approx.f64, ~2^-20) and two Newton steps (the second one
// brings the error to the last bit); ~40 cycles instead of ~300 on the loop-carried chains below
__device__ __forceinline__ double fast_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  y = fma(fma(e, e, e), y, y);
  return fma(fma(-d, y, 1.0), y, y);
}
__device__ __forceinline__ double damp_unscaled(double V, double s, double inv_radius) {
  const double s2 = s * s;
  return fmin(fmax(s2 * V, 1e-6), 1e32) * inv_radius * fast_rcp(s2);
}
// j-th (0-based) set bit of a <= 16-bit mask
__device__ __forceinline__ int nth_set_bit(unsigned mask, int j) {
  int pos = 0;
#pragma unroll
  for (int b = 0; b < 16; ++b) { if ((mask >> b) & 1u) { if (j == 0) pos = b; --j; } }
  return pos;
}

// -E^T W (lower triangle over the touched camera slots) and -E^T wg into the warp's accumulator; sE / sW hold DL rows of 64
template <int DL>
__device__ __forceinline__ void schur_contrib(double* acc, const double* sE, const double* sW, const int* slots, int t, int NT, int lane) {
  const int T6 = 6 * t;
  for (int ia = 0; ia < T6; ++ia) {   // slots ascend, so ib <= ia is the lower triangle
    const int I = 6 * slots[ia / 6] + ia % 6;
    for (int ib = lane; ib <= ia; ib += 32) {
      const int Jx = 6 * slots[ib / 6] + ib % 6;
      double v = 0.0;
#pragma unroll
      for (int d = 0; d < DL; ++d) v += sE[d * S_EC + I] * sW[d * S_EC + Jx];
      acc[I * (I + 1) / 2 + Jx] += v;
    }
  }
}

#define GSYNC() do { if (grid_sync(bar)) { if (threadIdx.x == 0) a.out[OUT_ABORT] = 1.0; return; } } while (0)

__global__ void __launch_bounds__(ST, 1) ba_small_kernel(SmallArgs a) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, n = a.n, NT = n * (n + 1) / 2, P = NT + n, LD = n + 1;
  double* s_cams[2] = {sm, sm + 7 * K};
  double* s_dc = sm + 14 * K;          // 64: camera step
  double* s_sc = s_dc + 64;            // 64: Jacobi scale of the camera columns
  double* s_misc = s_sc + 64;          // 32: broadcast slots
  double* s_red = s_misc + 32;         // 64
  double* s_il = s_red + 64;           // 64: 1 / sqrt(pivot) of the reduced system's columns
  double* s_cam3 = s_il + 64;          // 3 x 16: per free camera: gradient max-norm term, step^2, candidate norm^2
  int* s_slot = reinterpret_cast<int*>(s_cam3 + 48);   // K ints (K <= 256: 128 doubles)
  double* s_scr = s_cam3 + 48 + 128;   // SW x S_SCR
  double* s_acc = s_scr + SW * S_SCR;  // SW x P accumulators; reused as the (n + 1) x LD Cholesky workspace
  double* scr = s_scr + warp * S_SCR;
  double* acc = s_acc + warp * P;
  const int gw = blockIdx.x * SW + warp, GW = a.G * SW;
  GridBar bar{a.sync, a.sync + 1, a.G, 0};

  for (int e = tid; e < 7 * K; e += ST) s_cams[0][e] = a.cams_in[e];
  for (int e = tid; e < K; e += ST) s_slot[e] = a.camslot[e];
  for (int i = blockIdx.x * ST + tid; i < a.n_points; i += a.G * ST) a.rho[1][i] = a.rho[0][i];
  for (int i = blockIdx.x * ST + tid; i < 3 * a.n_planes; i += a.G * ST) a.theta[1][i] = a.theta[0][i];
  __syncthreads();
  {
    double v[3] = {0.0, 0.0, 0.0};
    eval_pass<true>(a, s_cams[0], a.rho[0], a.theta[0], a.rp[0], a.Jp[0], a.rt[0], a.Jt[0], v[0], v[1]);
    for (int l = blockIdx.x * ST + tid; l < a.nvp; l += a.G * ST) { const double x = a.rho[0][a.vp_gl[l]]; v[2] += x * x; }
    for (int l = blockIdx.x * ST + tid; l < 3 * a.nvt; l += a.G * ST) { const double x = a.theta[0][3 * a.vt_gl[l / 3] + l % 3]; v[2] += x * x; }
    block_sums<3>(v, s_red);
    if (tid == 0) { a.part0[4 * blockIdx.x] = v[0]; a.part0[4 * blockIdx.x + 1] = v[1]; a.part0[4 * blockIdx.x + 2] = v[2]; }
  }
  GSYNC();

  // ---- state: identical in every thread of every CTA ----
  double x_cost, fixed_cost, x_norm, radius = a.radius0, decrease_factor = 2.0;
  int iter = 0, n_ok = 0, n_bad = 0, invalid_run = 0, term = TSLAM_TERM_NO_CONVERGENCE, cur = 0;
  bool newJ = true, first = true;
  {
    if (warp < 3) { const double v = warp_sum_ctas(a.part0 + warp, 4, a.G, lane); if (lane == 0) s_misc[4 + warp] = v; }
    __syncthreads();
    if (tid == 0) {
      double xn = s_misc[6];
      for (int k = 0; k < K; ++k)
        if (s_slot[k] >= 0)
          for (int c = 0; c < 7; ++c) xn += s_cams[0][7 * k + c] * s_cams[0][7 * k + c];
      s_misc[0] = s_misc[4]; s_misc[1] = s_misc[5]; s_misc[2] = sqrt(xn);
    }
    __syncthreads();
    x_cost = s_misc[0]; fixed_cost = s_misc[1]; x_norm = s_misc[2];
    __syncthreads();
  }
  const double initial_cost = x_cost + fixed_cost;
  if (blockIdx.x == 0 && tid == 0) { double* t = a.out + OUT_HDR; t[0] = initial_cost; t[1] = radius; t[2] = 0.0; t[3] = 1.0; }

  while (true) {
    if (iter >= a.max_iters) { term = TSLAM_TERM_NO_CONVERGENCE; break; }
    if (radius <= 1e-32) { term = TSLAM_TERM_NO_CONVERGENCE; break; }
    if (a.nc + a.nvp + a.nvt == 0) { term = TSLAM_TERM_GRADIENT_TOL; break; }
    const double inv_radius = 1.0 / radius;
    const int pit = n_ok + n_bad;   // loop trip, for the phase stamps
    STAMP(0);
    const double *rp = a.rp[cur], *Jp = a.Jp[cur], *rt = a.rt[cur], *Jt = a.Jt[cur];

    // ---- camera block of the normal equations (only when the Jacobian is new) ----
    if (newJ && n > 0) {
      for (int e = tid; e < SW * P; e += ST) s_acc[e] = 0.0;
      __syncthreads();
      accum_pass<2, 13>(acc, scr, gw, a.lp, GW, a.p_cs, a.p_hs, Jp, rp, NT, lane);
      accum_pass<8, 15>(acc, scr, gw, a.lt, GW, a.t_cs, a.t_hs, Jt, rt, NT, lane);
      __syncthreads();
      for (int e = tid; e < P; e += ST) {
        double s = 0.0;
        for (int w = 0; w < SW; ++w) s += s_acc[w * P + e];
        a.partH[(size_t)blockIdx.x * P + e] = s;
      }
      __syncthreads();
    }
    STAMP(1);
    // ---- landmarks: V, g, E (new Jacobian), M^-1 for this radius, Schur complement contributions ----
    for (int e = tid; e < SW * P; e += ST) s_acc[e] = 0.0;
    __syncthreads();
    double gmax_l = 0.0, fail_l = 0.0;
    double* sE = scr + S_STG; double* sW = sE + 3 * S_EC; int* slots = reinterpret_cast<int*>(sW + 3 * S_EC);
    for (int v = gw; v < a.nvp; v += GW) {
      double e0 = 0.0, e1 = 0.0, V = 0.0, g = 0.0;
      unsigned mask = 0u;
      if (newJ) {
        const int s0 = lane / 6, k0 = lane % 6, s1 = (lane + 32) / 6, k1 = (lane + 32) % 6;
        const int o_end = a.vp_ptr[v + 1];
        for (int o0 = a.vp_ptr[v]; o0 < o_end; o0 += S_CHUNK) {
          const int cnt = min(S_CHUNK, o_end - o0);
          int oi = 0, ocs = -1, ohs = -1;
          if (lane < cnt) { oi = a.vp_obs[o0 + lane]; ocs = a.p_cs[oi]; ohs = a.p_hs[oi]; }
          __syncwarp();
          for (int q0 = 0; q0 < 28 * cnt; q0 += 32) {   // 26 Jacobian entries + 2 residuals per observation
            const int q = q0 + lane;
            const int k = min(q / 28, cnt - 1), e = q - 28 * k;
            const int i = __shfl_sync(0xffffffffu, oi, k);
            if (q < 28 * cnt) scr[q] = e < 26 ? Jp[26 * (size_t)i + e] : rp[2 * i + e - 26];
          }
          __syncwarp();
          for (int k = 0; k < cnt; ++k) {
            const int cs = __shfl_sync(0xffffffffu, ocs, k), hs = __shfl_sync(0xffffffffu, ohs, k);
            const double* Ji = scr + 28 * k;
            const double j0 = Ji[12], j1 = Ji[25];
            V += j0 * j0 + j1 * j1; g += j0 * Ji[26] + j1 * Ji[27];
            if (cs >= 0) mask |= 1u << cs;
            if (hs >= 0) mask |= 1u << hs;
            if (s0 == cs) e0 += j0 * Ji[k0] + j1 * Ji[13 + k0]; else if (s0 == hs) e0 += j0 * Ji[6 + k0] + j1 * Ji[19 + k0];
            if (s1 == cs) e1 += j0 * Ji[k1] + j1 * Ji[13 + k1]; else if (s1 == hs) e1 += j0 * Ji[6 + k1] + j1 * Ji[19 + k1];
          }
        }
        a.Ep[(size_t)v * S_EC + lane] = e0; a.Ep[(size_t)v * S_EC + lane + 32] = e1;
        if (lane == 0) { a.Vp[v] = V; a.gp[v] = g; a.maskp[v] = mask; if (first) a.sclp[v] = 1.0 / (1.0 + sqrt(V)); }
      } else {
        e0 = a.Ep[(size_t)v * S_EC + lane]; e1 = a.Ep[(size_t)v * S_EC + lane + 32];
        V = a.Vp[v]; g = a.gp[v]; mask = a.maskp[v];
      }
      const double sl = first ? 1.0 / (1.0 + sqrt(V)) : a.sclp[v];
      const double Minv = fast_rcp(V + damp_unscaled(V, sl, inv_radius));
      if (lane == 0) a.Mp[v] = Minv;
      gmax_l = fmax(gmax_l, fabs(g));
      if (n > 0 && mask) {
        const double wg = Minv * g;
        if (lane < n) acc[NT + lane] += e0 * wg;
        if (lane + 32 < n) acc[NT + lane + 32] += e1 * wg;
        sE[lane] = e0; sE[lane + 32] = e1; sW[lane] = Minv * e0; sW[lane + 32] = Minv * e1;
        const int t = __popc(mask);
        if (lane < t) slots[lane] = nth_set_bit(mask, lane);
        __syncwarp();
        schur_contrib<1>(acc, sE, sW, slots, t, NT, lane);
        __syncwarp();
      }
    }
    STAMP(14);
    // A plane belongs to a CTA (from the far end of the grid: the first CTAs carry the inverse depths): its observation list —
    // 25 features per keyframe that sees it — is dealt round-robin to the 8 warps, whose partial E / V / g are added in warp order by
    // warp 0, which then eliminates the plane. (One warp per list would pay an L2 round trip per block, 25+ in a row.)
    for (int v = a.G - 1 - (int)blockIdx.x; v < a.nvt; v += a.G) {
      double e[3][2] = {{0, 0}, {0, 0}, {0, 0}}, V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
      unsigned mask = 0u;
      if (newJ) {
        // 45 sums over the 8 rows of a block, two per lane: E (3 x [6 camera | 6 host] columns), V (6), g (3). A lane's E sums
        // belong to the columns of the block's camera / host slot; they are flushed into the E rows (shared memory) when that slot
        // changes along the list (blocks of one keyframe are consecutive, the host keyframe of a plane is usually one).
        int ia[2], ib[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int t = lane + 32 * h;
          if (t < 36) { ia[h] = 12 + t / 12; ib[h] = t % 12; }
          else if (t < 42) { const int k = t - 36; const int pa = k < 3 ? 0 : (k < 5 ? 1 : 2), pb = k < 3 ? k : (k < 5 ? k - 2 : 2); ia[h] = 12 + pa; ib[h] = 12 + pb; }
          else if (t < 45) { ia[h] = 12 + (t - 42); ib[h] = 15; }
          else { ia[h] = -1; ib[h] = 0; }
        }
        int offB[2], strB[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) { offB[h] = ib[h] < 15 ? ib[h] : 120; strB[h] = ib[h] < 15 ? 15 : 1; }
        for (int q = lane; q < 3 * S_EC; q += 32) sE[q] = 0.0;
        double ac[2] = {0.0, 0.0};
        int cur_cs = -2, cur_hs = -2;
        const int o_beg = a.vt_ptr[v], o_end = a.vt_ptr[v + 1];
        BlockRegs<8, 15> regs;
        __syncwarp();
        auto flush = [&](int slot, bool host_part) {   // add the lanes' E sums of the camera (ib < 6) or host (6 <= ib < 12) columns into row d of sE
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int t = lane + 32 * h;
            if (t < 36 && (ib[h] >= 6) == host_part) {
              if (slot >= 0) sE[(ia[h] - 12) * S_EC + 6 * slot + (host_part ? ib[h] - 6 : ib[h])] += ac[h];
              ac[h] = 0.0;
            }
          }
        };
        int o = o_beg + warp, oi = 0, cs = -1, hs = -1;
        if (o < o_end) { oi = a.vt_obs[o]; cs = a.t_cs[oi]; hs = a.t_hs[oi]; regs.load(Jt + 120 * (size_t)oi, rt + 8 * oi, lane); }
        while (o < o_end) {
          const int ccs = cs, chs = hs;
          regs.store(scr, lane);
          __syncwarp();
          o += SW;
          if (o < o_end) { oi = a.vt_obs[o]; cs = a.t_cs[oi]; hs = a.t_hs[oi]; regs.load(Jt + 120 * (size_t)oi, rt + 8 * oi, lane); }
          if (ccs != cur_cs) { flush(cur_cs, false); cur_cs = ccs; }
          if (chs != cur_hs) { flush(cur_hs, true); cur_hs = chs; }
          if (ccs >= 0) mask |= 1u << ccs;
          if (chs >= 0) mask |= 1u << chs;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (ia[h] < 0) continue;
            double sacc = 0.0;
#pragma unroll
            for (int row = 0; row < 8; ++row) sacc += scr[15 * row + ia[h]] * scr[offB[h] + strB[h] * row];
            ac[h] += sacc;
          }
          __syncwarp();
        }
        flush(cur_cs, false); flush(cur_hs, true);
        // this warp's partial V, g (lanes 4..12 of the second role) and slot mask beside its partial E rows
        if (lane >= 4 && lane < 13) sW[lane - 4] = ac[1];
        if (lane == 0) slots[0] = (int)mask;
        __syncthreads();
        if (warp == 0) {
          mask = 0u;
          for (int w = 0; w < SW; ++w) {
            const double* pE = s_scr + w * S_SCR + S_STG;
            const double* pW = pE + 3 * S_EC;
#pragma unroll
            for (int d = 0; d < 3; ++d) { e[d][0] += pE[d * S_EC + lane]; e[d][1] += pE[d * S_EC + lane + 32]; }
#pragma unroll
            for (int k = 0; k < 6; ++k) V[k] += pW[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) g[k] += pW[6 + k];
            mask |= (unsigned)reinterpret_cast<const int*>(pW + 3 * S_EC)[0];
          }
        }
        __syncthreads();
        if (warp != 0) continue;
#pragma unroll
        for (int d = 0; d < 3; ++d) { a.Et[((size_t)v * 3 + d) * S_EC + lane] = e[d][0]; a.Et[((size_t)v * 3 + d) * S_EC + lane + 32] = e[d][1]; }
        if (lane == 0) {
          for (int k = 0; k < 6; ++k) a.Vt[6 * (size_t)v + k] = V[k];
          for (int k = 0; k < 3; ++k) a.gt[3 * (size_t)v + k] = g[k];
          a.maskt[v] = mask;
          if (first) { a.sclt[3 * v] = 1.0 / (1.0 + sqrt(V[0])); a.sclt[3 * v + 1] = 1.0 / (1.0 + sqrt(V[3])); a.sclt[3 * v + 2] = 1.0 / (1.0 + sqrt(V[5])); }
        }
      } else {
        if (warp != 0) continue;
#pragma unroll
        for (int d = 0; d < 3; ++d) { e[d][0] = a.Et[((size_t)v * 3 + d) * S_EC + lane]; e[d][1] = a.Et[((size_t)v * 3 + d) * S_EC + lane + 32]; }
        for (int k = 0; k < 6; ++k) V[k] = a.Vt[6 * (size_t)v + k];
        for (int k = 0; k < 3; ++k) g[k] = a.gt[3 * (size_t)v + k];
        mask = a.maskt[v];
      }
      double sl[3];
      if (first) { sl[0] = 1.0 / (1.0 + sqrt(V[0])); sl[1] = 1.0 / (1.0 + sqrt(V[3])); sl[2] = 1.0 / (1.0 + sqrt(V[5])); }
      else { sl[0] = a.sclt[3 * v]; sl[1] = a.sclt[3 * v + 1]; sl[2] = a.sclt[3 * v + 2]; }
      // (V + D)^-1 by cofactors, as lm_vinv_kernel<3>
      const double ma = V[0] + damp_unscaled(V[0], sl[0], inv_radius), mb = V[1], mc = V[2];
      const double md = V[3] + damp_unscaled(V[3], sl[1], inv_radius), me = V[4], mf = V[5] + damp_unscaled(V[5], sl[2], inv_radius);
      const double c00 = md * mf - me * me, c01 = mc * me - mb * mf, c02 = mb * me - mc * md;
      const double det = ma * c00 + mb * c01 + mc * c02;
      if (!(det > 0.0) || !(ma > 0.0)) fail_l = 1.0;
      const double id = det > 0.0 ? fast_rcp(det) : 1.0 / det;
      const double Mi[6] = {c00 * id, c01 * id, c02 * id, (ma * mf - mc * mc) * id, (mb * mc - ma * me) * id, (ma * md - mb * mb) * id};
      if (lane == 0) for (int k = 0; k < 6; ++k) a.Mt[6 * (size_t)v + k] = Mi[k];
      gmax_l = fmax(gmax_l, fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2]))));
      if (n > 0 && mask) {
        const double wg[3] = {Mi[0] * g[0] + Mi[1] * g[1] + Mi[2] * g[2], Mi[1] * g[0] + Mi[3] * g[1] + Mi[4] * g[2], Mi[2] * g[0] + Mi[4] * g[1] + Mi[5] * g[2]};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = lane + 32 * h;
          const double w0 = Mi[0] * e[0][h] + Mi[1] * e[1][h] + Mi[2] * e[2][h];
          const double w1 = Mi[1] * e[0][h] + Mi[3] * e[1][h] + Mi[4] * e[2][h];
          const double w2 = Mi[2] * e[0][h] + Mi[4] * e[1][h] + Mi[5] * e[2][h];
          if (col < n) acc[NT + col] += e[0][h] * wg[0] + e[1][h] * wg[1] + e[2][h] * wg[2];
          sE[col] = e[0][h]; sE[S_EC + col] = e[1][h]; sE[2 * S_EC + col] = e[2][h];
          sW[col] = w0; sW[S_EC + col] = w1; sW[2 * S_EC + col] = w2;
        }
        const int t = __popc(mask);
        if (lane < t) slots[lane] = nth_set_bit(mask, lane);
        __syncwarp();
        schur_contrib<3>(acc, sE, sW, slots, t, NT, lane);
        __syncwarp();
      }
    }
    STAMP(15);
    __syncthreads();
    for (int e = tid; e < P; e += ST) {
      double s = 0.0;
      for (int w = 0; w < SW; ++w) s += s_acc[w * P + e];
      a.partS[(size_t)blockIdx.x * P + e] = s;
    }
    {
      // per-CTA maxima of the landmark gradient and of the 3x3 failure flag
      __syncthreads();   // (gmax_l / fail_l are warp-uniform: every lane walked the same observation list)
      if (lane == 0) { s_red[2 * warp] = gmax_l; s_red[2 * warp + 1] = fail_l; }
      __syncthreads();
      if (tid == 0) {
        double m0 = 0.0, m1 = 0.0;
        for (int w = 0; w < SW; ++w) { m0 = fmax(m0, s_red[2 * w]); m1 = fmax(m1, s_red[2 * w + 1]); }
        a.partG[2 * blockIdx.x] = m0; a.partG[2 * blockIdx.x + 1] = m1;
      }
    }
    STAMP(2);
    GSYNC();
    STAMP(3);
    // ---- fixed-order reduction over the CTAs, sliced across the grid ----
    for (int e = gw; e < P; e += GW) {   // a warp per entry: <= 5 independent loads per lane and array, then a fixed tree
      double sh = 0.0, ss = 0.0;
      if (newJ) for (int c = lane; c < a.G; c += 32) sh += a.partH[(size_t)c * P + e];
      for (int c = lane; c < a.G; c += 32) ss += a.partS[(size_t)c * P + e];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { sh += __shfl_xor_sync(0xffffffffu, sh, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
      if (lane == 0) { if (newJ) a.sumH[e] = sh; a.sumS[e] = ss; }
    }
    GSYNC();
    STAMP(4);
    // ---- reduced camera system: every CTA factors it (identical arithmetic, no broadcast) ----
    double* A = s_acc;   // (n + 1) x LD, lower triangle; row n carries the right-hand side
    if (first) for (int i = tid; i < n; i += ST) s_sc[i] = 1.0 / (1.0 + sqrt(a.sumH[i * (i + 1) / 2 + i]));
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    for (int i = ty; i < n; i += 16)
      for (int j = tx; j <= i; j += 16) {
        const int e = i * (i + 1) / 2 + j;
        const double h = a.sumH[e];
        double v = h - a.sumS[e];
        if (i == j) v += damp_unscaled(h, s_sc[i], inv_radius);
        A[i * LD + j] = v;
      }
    for (int j = tid; j < n; j += ST) A[n * LD + j] = a.sumH[NT + j] - a.sumS[NT + j];
    if (warp == 7) { const double v0 = warp_max_ctas(a.partG, 2, a.G, lane), v1 = warp_max_ctas(a.partG + 1, 2, a.G, lane); if (lane == 0) { s_misc[6] = v0; s_misc[7] = v1; } }
    __syncthreads();
    STAMP(5);
    // Right-looking Cholesky blocked by camera (6 columns): per block column (1) every thread factors the 6x6 diagonal block in
    // registers from broadcast shared-memory reads (no barrier, no hand-off: rsqrt.approx + a cubic step per pivot), (2) a thread per
    // row below solves its 6 entries against it (row n, the right-hand side, rides along and ends as y = L^-1 b), (3) the rank-6
    // update of the trailing rows on a 16 x 16 thread grid. Two barriers per camera instead of one per column (42 -> 14 on C4).
    int chol_fail = 0;
    for (int kb = 0; kb < a.nc; ++kb) {
      const int c0 = 6 * kb;
      double L[6][6], il[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double d = A[(c0 + c) * LD + c0 + c];
#pragma unroll
        for (int j = 0; j < c; ++j) d -= L[c][j] * L[c][j];
        if (!(d > 0.0)) chol_fail = 1;
        const double rs = rsqrt_pivot(d);
        il[c] = rs; L[c][c] = d * rs;
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
          double v = A[(c0 + r) * LD + c0 + c];
#pragma unroll
          for (int j = 0; j < c; ++j) v -= L[r][j] * L[c][j];
          L[r][c] = v * rs;
        }
      }
      if (chol_fail) break;   // uniform: every thread factored the same block
      __syncthreads();        // everybody has read the diagonal block
      if (tid < 6) {
#pragma unroll
        for (int c = 0; c < 6; ++c) if (tid == c) { s_il[c0 + c] = il[c]; for (int j = 0; j <= c; ++j) A[(c0 + c) * LD + c0 + j] = L[c][j]; }
      }
      for (int i = c0 + 6 + tid; i <= n; i += ST) {   // X L^T = A_panel: x_c = (a_c - sum_{j<c} x_j L_cj) / L_cc
        double x[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double v = A[i * LD + c0 + c];
#pragma unroll
          for (int j = 0; j < c; ++j) v -= x[j] * L[c][j];
          x[c] = v * il[c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) A[i * LD + c0 + c] = x[c];
      }
      __syncthreads();
      const int r0 = c0 + 6, na = (n - r0 + 16) >> 4;   // trailing rows r0 .. n
      for (int qa = 0; qa < na; ++qa) {
        const int i = r0 + ty + 16 * qa;
        if (i > n) break;
        double xi[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) xi[c] = A[i * LD + c0 + c];
        for (int qb = 0; qb <= qa; ++qb) {
          const int j = r0 + tx + 16 * qb;
          if (j > i || j >= n) continue;
          double acc = A[i * LD + j];
#pragma unroll
          for (int c = 0; c < 6; ++c) acc -= xi[c] * A[j * LD + c0 + c];
          A[i * LD + j] = acc;
        }
      }
      __syncthreads();
    }
    STAMP(6);
    if (!chol_fail && warp == 0) {
      // backward solve L^T u = y, columns in registers: lane holds u[lane], u[lane + 32]
      double y0 = lane < n ? A[n * LD + lane] : 0.0, y1 = lane + 32 < n ? A[n * LD + lane + 32] : 0.0;
      for (int k = n - 1; k >= 0; --k) {
        const double uk = __shfl_sync(0xffffffffu, k < 32 ? y0 : y1, k & 31) * s_il[k];
        if (lane == (k & 31)) { if (k < 32) y0 = uk; else y1 = uk; }
        if (lane < k) y0 -= A[k * LD + lane] * uk;
        if (lane + 32 < k) y1 -= A[k * LD + lane + 32] * uk;
      }
      if (lane < n) s_dc[lane] = -y0;
      if (lane + 32 < n) s_dc[lane + 32] = -y1;
    }
    __syncthreads();
    // candidate cameras, camera part of the step / candidate norms, gradient max-norm: a thread per camera, then slot order
    if (tid < K) {
      const int k = tid, sl = s_slot[k];
      const double* xk = s_cams[cur] + 7 * k;
      double* ok = s_cams[cur ^ 1] + 7 * k;
      if (sl < 0) { for (int c = 0; c < 7; ++c) ok[c] = xk[c]; }
      else {
        {   // ||x - Plus(x, -g)||_inf with the unscaled gradient
          const double* gr = a.sumH + NT + 6 * sl;
          const double d[3] = {-gr[0], -gr[1], -gr[2]};
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
          for (int c = 3; c < 6; ++c) m = fmax(m, fabs(gr[c]));
          s_cam3[sl] = m;
        }
        double step2 = 0.0, cn2 = 0.0;
        if (!chol_fail) {
          const double* d = s_dc + 6 * sl;
          const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
          double q[4];
          if (nrm > 0.0) {
            const double sn = sin(nrm) / nrm;
            const double z0 = cos(nrm), z1 = sn * d[0], z2 = sn * d[1], z3 = sn * d[2];
            q[0] = z0 * xk[0] - z1 * xk[1] - z2 * xk[2] - z3 * xk[3];
            q[1] = z0 * xk[1] + z1 * xk[0] + z2 * xk[3] - z3 * xk[2];
            q[2] = z0 * xk[2] - z1 * xk[3] + z2 * xk[0] + z3 * xk[1];
            q[3] = z0 * xk[3] + z1 * xk[2] - z2 * xk[1] + z3 * xk[0];
          } else { q[0] = xk[0]; q[1] = xk[1]; q[2] = xk[2]; q[3] = xk[3]; }
          for (int c = 0; c < 4; ++c) { ok[c] = q[c]; const double df = xk[c] - q[c]; step2 += df * df; cn2 += q[c] * q[c]; }
          for (int c = 0; c < 3; ++c) { const double v = xk[4 + c] + d[3 + c]; ok[4 + c] = v; step2 += d[3 + c] * d[3 + c]; cn2 += v * v; }
        }
        s_cam3[16 + sl] = step2; s_cam3[32 + sl] = cn2;
      }
    }
    __syncthreads();
    if (tid == 0) {
      double step2 = 0.0, cn2 = 0.0, gm = s_misc[6];
      for (int sl = 0; sl < a.nc; ++sl) { gm = fmax(gm, s_cam3[sl]); step2 += s_cam3[16 + sl]; cn2 += s_cam3[32 + sl]; }
      s_misc[0] = step2; s_misc[1] = cn2; s_misc[2] = gm; s_misc[3] = s_misc[7];
    }
    __syncthreads();
    STAMP(7);
    const double step2_c = s_misc[0], cn2_c = s_misc[1];
    const double gmax = s_misc[2];
    const bool lin_fail = chol_fail || s_misc[3] != 0.0;
    __syncthreads();
    if (gmax <= a.gtol) { term = TSLAM_TERM_GRADIENT_TOL; break; }   // Ceres tests the gradient before it computes a step
    ++iter;
    first = false;
    double sums[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // step^2 (landmarks), candidate norm^2 (landmarks), model cost change, candidate cost, candidate fixed cost
    if (!lin_fail) {
      // ---- landmark back-substitution and candidates ----
      double st2 = 0.0, c2 = 0.0;
      for (int v = gw; v < a.nvp; v += GW) {
        double t = (lane < n ? a.Ep[(size_t)v * S_EC + lane] * s_dc[lane] : 0.0) + (lane + 32 < n ? a.Ep[(size_t)v * S_EC + lane + 32] * s_dc[lane + 32] : 0.0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) {
          const double dl = -a.Mp[v] * (a.gp[v] + t);
          a.dlp[v] = dl;
          const int gl = a.vp_gl[v];
          const double xn = a.rho[cur][gl] + dl;
          a.rho[cur ^ 1][gl] = xn;
          st2 += dl * dl; c2 += xn * xn;
        }
      }
      for (int v = GW - 1 - gw; v < a.nvt; v += GW) {
        double t[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double* E = a.Et + ((size_t)v * 3 + d) * S_EC;
          t[d] = (lane < n ? E[lane] * s_dc[lane] : 0.0) + (lane + 32 < n ? E[lane + 32] * s_dc[lane + 32] : 0.0);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t[d] += __shfl_xor_sync(0xffffffffu, t[d], o);
        }
        if (lane == 0) {
          const double* Mi = a.Mt + 6 * (size_t)v;
          const double b0 = a.gt[3 * v] + t[0], b1 = a.gt[3 * v + 1] + t[1], b2 = a.gt[3 * v + 2] + t[2];
          const double dl[3] = {-(Mi[0] * b0 + Mi[1] * b1 + Mi[2] * b2), -(Mi[1] * b0 + Mi[3] * b1 + Mi[4] * b2), -(Mi[2] * b0 + Mi[4] * b1 + Mi[5] * b2)};
          const int gl = a.vt_gl[v];
          for (int d = 0; d < 3; ++d) {
            a.dlt[3 * v + d] = dl[d];
            const double xn = a.theta[cur][3 * gl + d] + dl[d];
            a.theta[cur ^ 1][3 * gl + d] = xn;
            st2 += dl[d] * dl[d]; c2 += xn * xn;
          }
        }
      }
      if (a.nvp + a.nvt > 0) {
        __syncthreads();
        if (lane == 0) { s_red[2 * warp] = st2; s_red[2 * warp + 1] = c2; }
        __syncthreads();
        if (tid == 0) { double s0 = 0.0, s1 = 0.0; for (int w = 0; w < SW; ++w) { s0 += s_red[2 * w]; s1 += s_red[2 * w + 1]; } s_misc[4] = s0; s_misc[5] = s1; }
        STAMP(8);
        GSYNC();
      } else if (tid == 0) { s_misc[4] = 0.0; s_misc[5] = 0.0; }
      STAMP(9);
      // ---- model cost change -(J d)'(r + J d / 2) and the candidate evaluation (with its Jacobian, speculatively) ----
      double v[3] = {0.0, 0.0, 0.0};
      for (int i = blockIdx.x * ST + tid; i < a.lp; i += a.G * ST) {
        const int cs = a.p_cs[i], hs = a.p_hs[i], ls = a.p_ls[i];
        if (cs < 0 && hs < 0 && ls < 0) continue;
        const double* Ji = Jp + 26 * (size_t)i;
        const double dl = ls >= 0 ? a.dlp[ls] : 0.0;
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          double m = Ji[13 * row + 12] * dl;
          if (cs >= 0) for (int k = 0; k < 6; ++k) m += Ji[13 * row + k] * s_dc[6 * cs + k];
          if (hs >= 0) for (int k = 0; k < 6; ++k) m += Ji[13 * row + 6 + k] * s_dc[6 * hs + k];
          v[0] -= m * (rp[2 * i + row] + m * 0.5);
        }
      }
      for (int gpx = blockIdx.x * ST + tid; gpx < 8 * a.lt; gpx += a.G * ST) {
        const int b = gpx >> 3;
        const int cs = a.t_cs[b], hs = a.t_hs[b], ls = a.t_ls[b];
        if (cs < 0 && hs < 0 && ls < 0) continue;
        const double* Jr = Jt + 15 * (size_t)gpx;
        double m = 0.0;
        if (ls >= 0) m = Jr[12] * a.dlt[3 * ls] + Jr[13] * a.dlt[3 * ls + 1] + Jr[14] * a.dlt[3 * ls + 2];
        if (cs >= 0) for (int k = 0; k < 6; ++k) m += Jr[k] * s_dc[6 * cs + k];
        if (hs >= 0) for (int k = 0; k < 6; ++k) m += Jr[6 + k] * s_dc[6 * hs + k];
        v[0] -= m * (rt[gpx] + m * 0.5);
      }
      STAMP(10);
      eval_pass<true>(a, s_cams[cur ^ 1], a.rho[cur ^ 1], a.theta[cur ^ 1], a.rp[cur ^ 1], a.Jp[cur ^ 1], a.rt[cur ^ 1], a.Jt[cur ^ 1], v[1], v[2]);
      block_sums<3>(v, s_red);
      if (tid == 0) {
        double* o = a.part + S_NPART * blockIdx.x;
        o[0] = s_misc[4]; o[1] = s_misc[5]; o[2] = v[0]; o[3] = v[1]; o[4] = v[2];
      }
      STAMP(11);
      GSYNC();
      STAMP(12);
      if (warp < 5) { const double v = warp_sum_ctas(a.part + warp, S_NPART, a.G, lane); if (lane == 0) s_misc[8 + warp] = v; }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 5; ++k) sums[k] = s_misc[8 + k];
      __syncthreads();
    }
    STAMP(13);
    // ---- accept / reject (run_lm of ba_solve.cu, statement for statement) ----
    const double mcc = sums[2], cand_cost = sums[3], step_norm = sqrt(step2_c + sums[0]);
    const bool finite = isfinite(mcc) && isfinite(cand_cost) && isfinite(step_norm);
    double* tr = a.out + OUT_HDR + 4 * iter;
    const bool writer = blockIdx.x == 0 && tid == 0;
    if (lin_fail || !finite || !(mcc > 0.0)) {
      ++invalid_run; ++n_bad;
      if (writer) { tr[0] = x_cost + fixed_cost; tr[1] = radius; tr[2] = 0.0; tr[3] = -1.0; }
      if (invalid_run >= 5) { term = TSLAM_TERM_FAILURE; break; }
      radius /= decrease_factor; decrease_factor *= 2.0;
      newJ = false;
      continue;
    }
    invalid_run = 0;
    if (step_norm <= a.ptol * (x_norm + a.ptol)) {
      term = TSLAM_TERM_PARAMETER_TOL;
      if (writer) { tr[0] = x_cost + fixed_cost; tr[1] = radius; tr[2] = 0.0; tr[3] = 0.0; }
      break;
    }
    const double cost_change = x_cost - cand_cost;
    if (fabs(cost_change) <= a.ftol * x_cost) {
      term = TSLAM_TERM_FUNCTION_TOL;
      if (writer) { tr[0] = x_cost + fixed_cost; tr[1] = radius; tr[2] = 0.0; tr[3] = 0.0; }
      break;
    }
    const double rel = cost_change / mcc;
    if (rel > 1e-3) {
      cur ^= 1;
      x_norm = sqrt(cn2_c + sums[1]);
      x_cost = cand_cost;
      const double q = 2.0 * rel - 1.0;
      radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - q * q * q));
      decrease_factor = 2.0;
      ++n_ok;
      newJ = true;
      if (writer) { tr[0] = x_cost + fixed_cost; tr[1] = radius; tr[2] = rel; tr[3] = 1.0; }
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0;
      ++n_bad;
      newJ = false;
      if (writer) { tr[0] = cand_cost + fixed_cost; tr[1] = radius; tr[2] = rel; tr[3] = 0.0; }
    }
  }
  // ---- results: summary + trace (CTA 0), parameters, Problem::Evaluate residuals at the solution ----
  __syncthreads();
  double* o_cams = a.out + OUT_HDR + 4 * (a.max_iters + 2);
  double* o_rho = o_cams + 7 * K;
  double* o_theta = o_rho + a.n_points;
  if (blockIdx.x == 0) {
    if (tid == 0) {
      a.out[OUT_ITER] = iter; a.out[OUT_OK] = n_ok; a.out[OUT_BAD] = n_bad; a.out[OUT_TERM] = term;
      a.out[OUT_INIT] = initial_cost; a.out[OUT_FINAL] = x_cost + fixed_cost; a.out[OUT_FIXED] = fixed_cost; a.out[OUT_ABORT] = 0.0;
    }
    for (int e = tid; e < 7 * K; e += ST) o_cams[e] = s_cams[cur][e];
  }
  for (int i = blockIdx.x * ST + tid; i < a.n_points; i += a.G * ST) o_rho[i] = a.rho[cur][i];
  for (int i = blockIdx.x * ST + tid; i < 3 * a.n_planes; i += a.G * ST) o_theta[i] = a.theta[cur][i];
  for (int i = blockIdx.x * ST + tid; i < 2 * a.lp; i += a.G * ST) a.r_final_p[i] = a.rp[cur][i];
  for (int i = blockIdx.x * ST + tid; i < 8 * a.lt; i += a.G * ST) a.r_final_t[i] = a.rt[cur][i];
}

// ------------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------------
struct SmallWorkspace {
  uint8_t* h_stage = nullptr; size_t h_cap = 0;     // page-locked: packed inputs out, result block back
  DevBuf<uint8_t> dev;                              // one device allocation, grow-only
  bool attr_set = false; size_t attr_smem = 0;
  ~SmallWorkspace() { if (h_stage) cudaFreeHost(h_stage); }
};

struct Packer {   // running offsets of 16-byte aligned segments
  size_t off = 0;
  size_t take(size_t bytes) { const size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; }
};

bool small_path_enabled() {
  const char* e = getenv("TSLAM_SMALL");   // read per call: the parity tests switch between the two paths inside one process
  return !(e && e[0] == '0');
}

void small_workspace_free(tslam_ctx* ctx) {
  if (ctx->small_ws) { delete static_cast<SmallWorkspace*>(ctx->small_ws); ctx->small_ws = nullptr; }
}

// Returns TSLAM_OK with *handled = false when the problem is not eligible (the caller takes the general path).
int small_solve(tslam_ctx* ctx, tslam_ba_problem* p, const tslam_solve_options* opt, tslam_solve_summary* summary, double* final_residuals,
                double* trace, const double** d_rp, const double** d_rt, bool* handled) {
  *handled = false;
  if (!small_path_enabled() || ctx->world > 1) return TSLAM_OK;
  if (opt->text_jac_mode != TSLAM_JAC_ANALYTIC && p->n_tobs > 0) return TSLAM_OK;
  const int K = p->n_cams, lp = p->n_pobs, lt = p->n_tobs, NP = p->n_points, NPL = p->n_planes;
  if (K > S_MAX_K || lp > S_MAX_P || lt > S_MAX_T || lp + lt == 0 || opt->max_iters > 1000) return TSLAM_OK;
  auto T0 = std::chrono::steady_clock::now();
  // ---- structure: free-camera slots, free landmarks and their observation lists (one pass each) ----
  std::vector<int> camslot(K, -1), lsP(NP, -1), lsT(NPL, -1), cntP, cntT;
  std::vector<uint8_t> used(K, 0);
  auto cfix = [&](int k) { return p->cam_fixed && p->cam_fixed[k]; };
  for (int i = 0; i < lp; ++i) {
    const int l = p->p_lm[i];
    const bool lfree = !(p->rho_fixed && p->rho_fixed[l]);
    if (lfree || !cfix(p->p_cam[i]) || !cfix(p->p_host[i])) { used[p->p_cam[i]] = 1; used[p->p_host[i]] = 1; }
    if (lfree) { if (lsP[l] < 0) { lsP[l] = (int)cntP.size(); cntP.push_back(0); } ++cntP[lsP[l]]; }
  }
  for (int i = 0; i < lt; ++i) {
    const int l = p->t_plane[i];
    const bool lfree = !(p->theta_fixed && p->theta_fixed[l]);
    if (lfree || !cfix(p->t_cam[i]) || !cfix(p->t_host[i])) { used[p->t_cam[i]] = 1; used[p->t_host[i]] = 1; }
    if (lfree) { if (lsT[l] < 0) { lsT[l] = (int)cntT.size(); cntT.push_back(0); } ++cntT[lsT[l]]; }
  }
  int nc = 0;
  for (int k = 0; k < K; ++k) if (!cfix(k) && used[k]) camslot[k] = nc++;
  if (nc > S_MAX_NC) return TSLAM_OK;
  const int nvp = (int)cntP.size(), nvt = (int)cntT.size();
  for (int c : cntP) if (c > S_MAX_LIST) return TSLAM_OK;
  for (int c : cntT) if (c > S_MAX_LIST) return TSLAM_OK;
  const int n = 6 * nc, NT = n * (n + 1) / 2, P = NT + n;
  // one CTA per SM as soon as there is work for it: the per-warp lists (observations to accumulate, landmarks to eliminate) are what
  // an iteration waits for, and they shrink with the number of warps
  int G = (lp + lt + nvp + nvt + 7) / 8;
  G = std::max(1, std::min(G, std::min(S_MAX_G, ctx->sm_count)));
  if (const char* e = getenv("TSLAM_SMALL_G")) G = std::max(1, std::min(atoi(e), std::min(S_MAX_G, ctx->sm_count)));   // tuning experiments
  const int max_iters = opt->max_iters;

  // ---- layout ----
  Packer in;   // uploaded block
  const size_t o_sync = in.take(64);
  const size_t o_cams = in.take(sizeof(double) * 7 * K), o_rho = in.take(sizeof(double) * NP), o_theta = in.take(sizeof(double) * 3 * NPL);
  const size_t o_puv = in.take(sizeof(double) * 2 * lp), o_pray = in.take(sizeof(double) * 2 * lp);
  const size_t o_trays = in.take(sizeof(double) * 16 * lt), o_tiref = in.take(sizeof(double) * 8 * lt), o_tms = in.take(sizeof(double) * 2 * lt);
  const size_t o_pcam = in.take(4 * (size_t)lp), o_phost = in.take(4 * (size_t)lp), o_plm = in.take(4 * (size_t)lp), o_pls = in.take(4 * (size_t)lp);
  const size_t o_pcs = in.take(4 * (size_t)lp), o_phs = in.take(4 * (size_t)lp), o_tcs = in.take(4 * (size_t)lt), o_ths = in.take(4 * (size_t)lt);
  const size_t o_tcam = in.take(4 * (size_t)lt), o_thost = in.take(4 * (size_t)lt), o_tpl = in.take(4 * (size_t)lt), o_timg = in.take(4 * (size_t)lt), o_tls = in.take(4 * (size_t)lt);
  const size_t o_slot = in.take(4 * (size_t)K);
  const size_t o_vpptr = in.take(4 * (size_t)(nvp + 1)), o_vpobs = in.take(4 * (size_t)lp), o_vpgl = in.take(4 * (size_t)nvp);
  const size_t o_vtptr = in.take(4 * (size_t)(nvt + 1)), o_vtobs = in.take(4 * (size_t)lt), o_vtgl = in.take(4 * (size_t)nvt);
  const size_t in_bytes = in.off;
  Packer dv; dv.off = in_bytes;   // device-only segments behind the mirror of the uploaded block
  const size_t img_bytes = lt ? (size_t)p->n_imgs * p->img_w * p->img_h : 0;
  const size_t o_img = dv.take(img_bytes);
  const size_t o_rho1 = dv.take(sizeof(double) * NP), o_theta1 = dv.take(sizeof(double) * 3 * NPL);
  size_t o_rp[2], o_Jp[2], o_rt[2], o_Jt[2];
  for (int s = 0; s < 2; ++s) { o_rp[s] = dv.take(sizeof(double) * 2 * lp); o_Jp[s] = dv.take(sizeof(double) * 26 * lp); o_rt[s] = dv.take(sizeof(double) * 8 * lt); o_Jt[s] = dv.take(sizeof(double) * 120 * lt); }
  const size_t o_Vp = dv.take(8 * (size_t)nvp), o_gp = dv.take(8 * (size_t)nvp), o_Ep = dv.take(8 * (size_t)nvp * S_EC), o_Mp = dv.take(8 * (size_t)nvp), o_sclp = dv.take(8 * (size_t)nvp), o_dlp = dv.take(8 * (size_t)nvp);
  const size_t o_Vt = dv.take(8 * (size_t)nvt * 6), o_gt = dv.take(8 * (size_t)nvt * 3), o_Et = dv.take(8 * (size_t)nvt * 3 * S_EC), o_Mt = dv.take(8 * (size_t)nvt * 6), o_sclt = dv.take(8 * (size_t)nvt * 3), o_dlt = dv.take(8 * (size_t)nvt * 3);
  const size_t o_maskp = dv.take(4 * (size_t)nvp), o_maskt = dv.take(4 * (size_t)nvt);
  const size_t o_partH = dv.take(8 * (size_t)G * P), o_partS = dv.take(8 * (size_t)G * P), o_sumH = dv.take(8 * (size_t)P), o_sumS = dv.take(8 * (size_t)P);
  const size_t o_part0 = dv.take(8 * (size_t)G * 4), o_partG = dv.take(8 * (size_t)G * 2), o_part = dv.take(8 * (size_t)G * S_NPART);
  // result block: [summary + trace + parameters | point residuals | text residuals], three 16-byte aligned segments copied back at once
  // (the chi^2 gate kernels read the residual segments with vector loads)
  const size_t par_doubles = OUT_HDR + 4 * (size_t)(max_iters + 2) + 7 * (size_t)K + NP + 3 * (size_t)NPL;
  const size_t o_out = dv.take(8 * par_doubles), o_frp = dv.take(8 * 2 * (size_t)lp), o_frt = dv.take(8 * 8 * (size_t)lt);
  const size_t out_bytes = dv.off - o_out;
  const bool prof = getenv("TSLAM_SMALL_PROF") != nullptr;
  const size_t o_prof = dv.take(prof ? 8 * 16 * 16 : 0);
  const size_t dev_bytes = dv.off;

  if (!ctx->small_ws) ctx->small_ws = new SmallWorkspace();
  SmallWorkspace& W = *static_cast<SmallWorkspace*>(ctx->small_ws);
  const size_t stage_need = std::max(in_bytes, out_bytes);
  if (W.h_cap < stage_need) {
    if (W.h_stage) { TSL_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(W.h_stage); W.h_stage = nullptr; W.h_cap = 0; }
    const size_t cap = stage_need + stage_need / 2;
    TSL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&W.h_stage), cap, cudaHostAllocDefault));
    W.h_cap = cap;
  }
  TSL_CUDA(W.dev.reserve(dev_bytes + dev_bytes / 4));

  // ---- pack ----
  uint8_t* hs = W.h_stage;
  memset(hs + o_sync, 0, 64);
  auto putd = [&](size_t off, const double* src, size_t cnt) { if (cnt) memcpy(hs + off, src, sizeof(double) * cnt); };
  auto puti = [&](size_t off, const int32_t* src, size_t cnt) { if (cnt) memcpy(hs + off, src, 4 * cnt); };
  putd(o_cams, p->cams, 7 * (size_t)K); putd(o_rho, p->rho, NP); putd(o_theta, p->theta, 3 * (size_t)NPL);
  putd(o_puv, p->p_uv, 2 * (size_t)lp); putd(o_pray, p->p_ray, 2 * (size_t)lp);
  putd(o_trays, p->t_rays, 16 * (size_t)lt); putd(o_tiref, p->t_iref, 8 * (size_t)lt); putd(o_tms, p->t_musigma, 2 * (size_t)lt);
  puti(o_pcam, p->p_cam, lp); puti(o_phost, p->p_host, lp); puti(o_plm, p->p_lm, lp);
  puti(o_tcam, p->t_cam, lt); puti(o_thost, p->t_host, lt); puti(o_tpl, p->t_plane, lt); puti(o_timg, p->t_img, lt);
  puti(o_slot, camslot.data(), K);
  {
    int32_t *pcs = reinterpret_cast<int32_t*>(hs + o_pcs), *phs = reinterpret_cast<int32_t*>(hs + o_phs), *tcs = reinterpret_cast<int32_t*>(hs + o_tcs), *ths = reinterpret_cast<int32_t*>(hs + o_ths);
    for (int i = 0; i < lp; ++i) { pcs[i] = camslot[p->p_cam[i]]; phs[i] = camslot[p->p_host[i]]; }
    for (int i = 0; i < lt; ++i) { tcs[i] = camslot[p->t_cam[i]]; ths[i] = camslot[p->t_host[i]]; }
    int32_t* pls = reinterpret_cast<int32_t*>(hs + o_pls);
    int32_t* vptr = reinterpret_cast<int32_t*>(hs + o_vpptr); int32_t* vobs = reinterpret_cast<int32_t*>(hs + o_vpobs); int32_t* vgl = reinterpret_cast<int32_t*>(hs + o_vpgl);
    vptr[0] = 0;
    for (int v = 0; v < nvp; ++v) vptr[v + 1] = vptr[v] + cntP[v];
    std::vector<int> fill(vptr, vptr + nvp);
    for (int i = 0; i < lp; ++i) { const int s = lsP[p->p_lm[i]]; pls[i] = s; if (s >= 0) { vobs[fill[s]++] = i; vgl[s] = p->p_lm[i]; } }
    int32_t* tls = reinterpret_cast<int32_t*>(hs + o_tls);
    int32_t* tptr = reinterpret_cast<int32_t*>(hs + o_vtptr); int32_t* tobs = reinterpret_cast<int32_t*>(hs + o_vtobs); int32_t* tgl = reinterpret_cast<int32_t*>(hs + o_vtgl);
    tptr[0] = 0;
    for (int v = 0; v < nvt; ++v) tptr[v + 1] = tptr[v] + cntT[v];
    std::vector<int> fillt(tptr, tptr + nvt);
    for (int i = 0; i < lt; ++i) { const int s = lsT[p->t_plane[i]]; tls[i] = s; if (s >= 0) { tobs[fillt[s]++] = i; tgl[s] = p->t_plane[i]; } }
  }
  auto T1 = std::chrono::steady_clock::now();
  cudaStream_t st = ctx->stream;
  uint8_t* db = W.dev.p;
  TSL_CUDA(cudaMemcpyAsync(db, hs, in_bytes, cudaMemcpyHostToDevice, st));
  if (img_bytes) TSL_CUDA(cudaMemcpyAsync(db + o_img, p->imgs, img_bytes, cudaMemcpyHostToDevice, st));

  SmallArgs a{};
  auto D = [&](size_t off) { return reinterpret_cast<double*>(db + off); };
  auto I = [&](size_t off) { return reinterpret_cast<int*>(db + off); };
  a.p_uv = reinterpret_cast<const double2*>(db + o_puv); a.p_ray = reinterpret_cast<const double2*>(db + o_pray);
  a.p_cam = I(o_pcam); a.p_host = I(o_phost); a.p_lm = I(o_plm); a.p_ls = I(o_pls); a.p_cs = I(o_pcs); a.p_hs = I(o_phs); a.t_cs = I(o_tcs); a.t_hs = I(o_ths);
  a.t_rays = reinterpret_cast<const double2*>(db + o_trays); a.t_musigma = reinterpret_cast<const double2*>(db + o_tms); a.t_iref = D(o_tiref);
  a.t_cam = I(o_tcam); a.t_host = I(o_thost); a.t_plane = I(o_tpl); a.t_img = I(o_timg); a.t_ls = I(o_tls);
  a.imgs = db + o_img; a.img_w = p->img_w; a.img_h = p->img_h;
  a.camslot = I(o_slot);
  a.vp_ptr = I(o_vpptr); a.vp_obs = I(o_vpobs); a.vp_gl = I(o_vpgl); a.vt_ptr = I(o_vtptr); a.vt_obs = I(o_vtobs); a.vt_gl = I(o_vtgl);
  a.pfx = p->K_point[0]; a.pfy = p->K_point[1]; a.pcx = p->K_point[2]; a.pcy = p->K_point[3]; a.wx = p->w_point[0]; a.wy = p->w_point[1]; a.hub_p = p->huber_point;
  a.tfx = p->K_text[0]; a.tfy = p->K_text[1]; a.tcx = p->K_text[2]; a.tcy = p->K_text[3]; a.wT = p->w_text; a.hub_t = p->huber_text;
  a.cams_in = D(o_cams);
  a.rho[0] = D(o_rho); a.rho[1] = D(o_rho1); a.theta[0] = D(o_theta); a.theta[1] = D(o_theta1);
  for (int s = 0; s < 2; ++s) { a.rp[s] = D(o_rp[s]); a.Jp[s] = D(o_Jp[s]); a.rt[s] = D(o_rt[s]); a.Jt[s] = D(o_Jt[s]); }
  a.Vp = D(o_Vp); a.gp = D(o_gp); a.Ep = D(o_Ep); a.Mp = D(o_Mp); a.sclp = D(o_sclp); a.dlp = D(o_dlp);
  a.Vt = D(o_Vt); a.gt = D(o_gt); a.Et = D(o_Et); a.Mt = D(o_Mt); a.sclt = D(o_sclt); a.dlt = D(o_dlt);
  a.maskp = reinterpret_cast<unsigned*>(db + o_maskp); a.maskt = reinterpret_cast<unsigned*>(db + o_maskt);
  a.partH = D(o_partH); a.partS = D(o_partS); a.sumH = D(o_sumH); a.sumS = D(o_sumS); a.part0 = D(o_part0); a.partG = D(o_partG); a.part = D(o_part);
  a.sync = I(o_sync);
  a.out = D(o_out);
  a.prof = prof ? reinterpret_cast<long long*>(db + o_prof) : nullptr;
  if (prof) TSL_CUDA(cudaMemsetAsync(db + o_prof, 0, 8 * 16 * 16, st));
  a.r_final_p = D(o_frp); a.r_final_t = D(o_frt);
  a.K = K; a.nc = nc; a.n = n; a.lp = lp; a.lt = lt; a.nvp = nvp; a.nvt = nvt; a.n_points = NP; a.n_planes = NPL; a.G = G; a.max_iters = max_iters;
  a.ftol = opt->function_tolerance > 0 ? opt->function_tolerance : 1e-6;
  a.gtol = opt->gradient_tolerance > 0 ? opt->gradient_tolerance : 1e-10;
  a.ptol = opt->parameter_tolerance > 0 ? opt->parameter_tolerance : 1e-8;
  a.radius0 = opt->initial_radius > 0 ? opt->initial_radius : 1e4;

  const size_t acc_doubles = std::max((size_t)SW * P, (size_t)(n + 1) * (n + 1));
  const size_t smem = sizeof(double) * (14 * (size_t)K + 64 + 64 + 32 + 64 + 64 + 48 + 128 + (size_t)SW * S_SCR + acc_doubles);
  if (!W.attr_set || W.attr_smem < smem) {
    TSL_CUDA(cudaFuncSetAttribute(ba_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    W.attr_set = true; W.attr_smem = 220 * 1024;
  }
  if (smem > 220 * 1024) return TSLAM_OK;
  {   // the grid barrier needs every CTA resident: cap the grid by what the device can hold with this much shared memory (a smaller
      // carve-out, MPS limits, a partitioned GPU); nothing resident -> not eligible, the general path takes the problem
    int per_sm = 0;
    TSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ba_small_kernel, ST, smem));
    if (per_sm < 1) return TSLAM_OK;
    if (G > per_sm * ctx->sm_count) { G = per_sm * ctx->sm_count; a.G = G; }
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(ST); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;   // co-residency of the G CTAs is what the grid barrier needs
    cfg.attrs = attr; cfg.numAttrs = 1;
    ++g_launches;
    TSL_CUDA(cudaLaunchKernelEx(&cfg, ba_small_kernel, a));
  }
  TSL_CUDA(cudaMemcpyAsync(hs, db + o_out, out_bytes, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  auto T2 = std::chrono::steady_clock::now();
  const double* out = reinterpret_cast<const double*>(hs);
  if (prof) {   // phase laps of CTA 0 in SM clock cycles (stamps: see STAMP() in the kernel)
    std::vector<long long> pr(16 * 16);
    TSL_CUDA(cudaMemcpy(pr.data(), db + o_prof, 8 * 16 * 16, cudaMemcpyDeviceToHost));
    static const int order[15] = {1, 14, 15, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13};
    static const char* names[15] = {"accumH", "zero+points", "planes", "partS", "bar2", "slice+bar3", "build", "chol", "solve+cams", "backsub", "bar4", "mcc", "eval", "bar5", "sums"};
    for (int it = 0; it < 16 && pr[16 * it]; ++it) {
      fprintf(stderr, "[small prof] trip %d (G=%d):", it, G);
      int prev = 0;
      for (int k = 0; k < 15; ++k) if (pr[16 * it + order[k]]) { fprintf(stderr, " %s %lld", names[k], pr[16 * it + order[k]] - pr[16 * it + prev]); prev = order[k]; }
      fprintf(stderr, "\n");
    }
  }
  tslam_solve_summary sum{};
  sum.iterations = (int)out[OUT_ITER]; sum.successful_steps = (int)out[OUT_OK]; sum.unsuccessful_steps = (int)out[OUT_BAD]; sum.termination = (int)out[OUT_TERM];
  sum.initial_cost = out[OUT_INIT]; sum.final_cost = out[OUT_FINAL]; sum.fixed_cost = out[OUT_FIXED];
  sum.n_free_cams = nc; sum.n_free_points = nvp; sum.n_free_planes = nvt; sum.reduced_dim = n;
  const double* o_tr = out + OUT_HDR;
  if (trace) memcpy(trace, o_tr, sizeof(double) * 4 * (size_t)(sum.iterations + 1));
  const double* oc = o_tr + 4 * (size_t)(max_iters + 2);
  memcpy(p->cams, oc, sizeof(double) * 7 * (size_t)K);
  if (NP) memcpy(p->rho, oc + 7 * (size_t)K, sizeof(double) * NP);
  if (NPL) memcpy(p->theta, oc + 7 * (size_t)K + NP, sizeof(double) * 3 * (size_t)NPL);
  if (final_residuals) {
    if (lp) memcpy(final_residuals, hs + (o_frp - o_out), sizeof(double) * 2 * (size_t)lp);
    if (lt) memcpy(final_residuals + 2 * (size_t)lp, hs + (o_frt - o_out), sizeof(double) * 8 * (size_t)lt);
  }
  if (d_rp) *d_rp = a.r_final_p;
  if (d_rt) *d_rt = a.r_final_t;
  auto T3 = std::chrono::steady_clock::now();
  sum.setup_ms = std::chrono::duration<double, std::milli>(T1 - T0).count();
  sum.solve_ms = std::chrono::duration<double, std::milli>(T2 - T1).count();
  sum.total_ms = std::chrono::duration<double, std::milli>(T3 - T0).count();
  if (summary) *summary = sum;
  *handled = true;
  return TSLAM_OK;
}

}  // namespace tsl
