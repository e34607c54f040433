// Residual + Jacobian evaluation kernels (the metric kernels of BASELINE.json) for sm_100a.
//
// Replaces, per residual block, ceres::AutoDiffCostFunction<auto_BAScene,2,4,3,4,3,1>::Evaluate
// (include/auto_BAScene.h:89-92 and the NW / PoseOptim / Rho variants) and
// ceres::NumericDiffCostFunction<nume_BAText,CENTRAL,8,4,3,4,3,3>::Evaluate
// (include/nume_BAText.h:97-100 and the PoseOptim / theta variants), followed by the
// QuaternionParameterization projection Ceres' evaluator applies.
//
// HBM layout / traffic: inputs are observation-major SoA streams read once with
// 16-byte (double2) coalesced loads; camera / landmark parameter blocks are tiny and L1/L2 resident
// (gathered through the read-only path); outputs are observation-major r[N][2|8] and
// J[N][rows][ncols], staged through shared memory so that every warp store instruction writes
// full, contiguous 32-byte sectors. Algorithmic bytes: 268 B per point BA eval, 1280 B per text BA
// block.
#include "ctx.cuh"
#include "ba_device.cuh"

namespace tsl {

constexpr int kEvalThreads = 128;

struct PointArgs {
  const double* cams; const double* rho;
  const double2* uv; const double2* ray;
  const int32_t* cam; const int32_t* host; const int32_t* lm;
  double fx, fy, cx, cy, wx, wy;
  int n;
  // robust (solver) mode
  double huber; const uint8_t* active; double* cost_part;  // cost_part[2*grid]: active, fixed
};

// COL0 = first tangent column exported, NCOLS = number of columns (13: BA, 6: pose, 1 @12: rho)
template <int COL0, int NCOLS, bool WANT_J, bool ROBUST, int MINB = 6>
__global__ void __launch_bounds__(kEvalThreads, MINB) point_eval_kernel(PointArgs a, double2* __restrict__ r_out, double* __restrict__ J_out) {
  PDL_PROLOGUE();
  constexpr int ROW = 2 * NCOLS;
  constexpr int STRIDE = (ROW % 2 == 0) ? ROW + 1 : ROW;  // odd stride in doubles: conflict-free 64-bit smem access
  __shared__ double sJ[WANT_J ? kEvalThreads * STRIDE : 1];
  __shared__ double sred[ROBUST ? 2 * (kEvalThreads / 32) : 1];
  const int i = blockIdx.x * kEvalThreads + threadIdx.x;
  double cost_a = 0.0, cost_f = 0.0;
  if (i < a.n) {
    const double2 uv = a.uv[i];
    const double2 ray = a.ray[i];
    const int ci = __ldg(a.cam + i), hi = __ldg(a.host + i), li = __ldg(a.lm + i);
    const Cam c = load_cam(a.cams, ci);
    const Cam h = load_cam(a.cams, hi);
    const double rho = __ldg(a.rho + li);
    double r[2], J[26];
    point_eval<WANT_J>(c, h, rho, ray.x, ray.y, uv.x, uv.y, a.fx, a.fy, a.cx, a.cy, a.wx, a.wy, r, J);
    double sq = 1.0;
    if (ROBUST) {
      double rho0;
      sq = huber_scale(a.huber, r[0] * r[0] + r[1] * r[1], &rho0);
      const bool act = a.active ? a.active[i] != 0 : true;
      if (act) cost_a = 0.5 * rho0; else cost_f = 0.5 * rho0;
      r[0] *= sq; r[1] *= sq;
    }
    r_out[i] = make_double2(r[0], r[1]);
    if (WANT_J) {
      double* s = sJ + threadIdx.x * STRIDE;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int cidx = 0; cidx < NCOLS; ++cidx) s[k * NCOLS + cidx] = J[13 * k + COL0 + cidx] * sq;
    }
  }
  if (WANT_J) {
    __syncthreads();
    const int base_row = blockIdx.x * kEvalThreads;
    const int nvalid = min(kEvalThreads, a.n - base_row);
    const int total2 = nvalid * ROW / 2;  // number of double2 elements (ROW is even)
    double2* out2 = reinterpret_cast<double2*>(J_out + (size_t)base_row * ROW);
    for (int e = threadIdx.x; e < total2; e += kEvalThreads) {
      const int idx = 2 * e;
      const int row = idx / ROW, col = idx - row * ROW;
      const double* s = sJ + row * STRIDE + col;
      out2[e] = make_double2(s[0], s[1]);
    }
  }
  if (ROBUST) {
    // deterministic block reduction: warp shuffle tree, then fixed-order sum over warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cost_a += __shfl_down_sync(0xffffffffu, cost_a, o);
      cost_f += __shfl_down_sync(0xffffffffu, cost_f, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sred[2 * w] = cost_a; sred[2 * w + 1] = cost_f; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double sa = 0, sf = 0;
      for (int k = 0; k < kEvalThreads / 32; ++k) { sa += sred[2 * k]; sf += sred[2 * k + 1]; }
      a.cost_part[2 * blockIdx.x] = sa;
      a.cost_part[2 * blockIdx.x + 1] = sf;
    }
  }
}

struct TextArgs {
  const double* cams; const double* theta;
  const double2* rays;   // n x 8
  const double* iref;    // n x 8
  const double2* musigma;
  const int32_t* cam; const int32_t* host; const int32_t* plane; const int32_t* img;
  const uint8_t* imgs; int img_w, img_h;
  double fx, fy, cx, cy, wT;
  int n;  // number of text blocks
  unsigned free_mask; const uint8_t* free_masks;  // per-block mask overrides free_mask when non-null
  double huber; const uint8_t* active; double* cost_part;
};

// One lane per pattern pixel: 4 text blocks per warp, 16 per CTA.
template <int COL0, int NCOLS, int MODE, bool WANT_J, bool ROBUST>
__global__ void __launch_bounds__(kEvalThreads) text_eval_kernel(TextArgs a, double* __restrict__ r_out, double* __restrict__ J_out) {
  PDL_PROLOGUE();
  constexpr int STRIDE = (NCOLS % 2 == 0) ? NCOLS + 1 : NCOLS;
  __shared__ double sJ[WANT_J ? kEvalThreads * STRIDE : 1];
  __shared__ double sred[ROBUST ? 2 * (kEvalThreads / 32) : 1];
  const int gpx = blockIdx.x * kEvalThreads + threadIdx.x;  // global pixel-row index
  const int b = gpx >> 3, px = gpx & 7;
  double cost_a = 0.0, cost_f = 0.0;
  const bool valid = b < a.n;
  double res = 0.0, Jr[15];
  if (valid) {
    const Cam c = load_cam(a.cams, __ldg(a.cam + b));
    const Cam h = load_cam(a.cams, __ldg(a.host + b));
    const double* thp = a.theta + 3 * (size_t)__ldg(a.plane + b);
    const double th[3] = {__ldg(thp), __ldg(thp + 1), __ldg(thp + 2)};
    const double2 ray = a.rays[gpx];
    const double iref = a.iref[gpx];
    const double2 ms = a.musigma[b];
    TextImg im{a.imgs + (size_t)__ldg(a.img + b) * a.img_w * a.img_h, a.img_w, a.img_h};
    if (!WANT_J) {
      res = text_residual_only(c.q, c.t, h.q, h.t, th, ray.x, ray.y, im, a.fx, a.fy, a.cx, a.cy, ms.x, ms.y, iref, a.wT);
    } else if (MODE == TSLAM_JAC_CENTRAL_DIFF) {
      const unsigned m = a.free_masks ? a.free_masks[b] : a.free_mask;
      res = text_pixel_central(c, h, th, ray.x, ray.y, im, a.fx, a.fy, a.cx, a.cy, ms.x, ms.y, iref, a.wT, m, Jr);
    } else {
      res = text_pixel_analytic(c, h, th, ray.x, ray.y, im, a.fx, a.fy, a.cx, a.cy, ms.x, ms.y, iref, a.wT, Jr);
    }
  }
  double sq = 1.0;
  if (ROBUST) {
    // s = ||r_block||^2 over the 8 lanes of this text block (xor butterfly -> identical in all 8 lanes)
    double s = res * res;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    double rho0;
    sq = huber_scale(a.huber, s, &rho0);
    if (valid && px == 0) {
      const bool act = a.active ? a.active[b] != 0 : true;
      if (act) cost_a = 0.5 * rho0; else cost_f = 0.5 * rho0;
    }
    res *= sq;
  }
  if (valid) r_out[gpx] = res;
  if (WANT_J) {
    if (valid) {
      double* s = sJ + threadIdx.x * STRIDE;
#pragma unroll
      for (int k = 0; k < NCOLS; ++k) s[k] = Jr[COL0 + k] * sq;
    }
    __syncthreads();
    const int base_row = blockIdx.x * kEvalThreads;
    const int nvalid = min(kEvalThreads, 8 * a.n - base_row);
    const int total = nvalid * NCOLS;
    double* out = J_out + (size_t)base_row * NCOLS;
    for (int e = threadIdx.x; e < total; e += kEvalThreads) {
      const int row = e / NCOLS, col = e - row * NCOLS;
      out[e] = sJ[row * STRIDE + col];
    }
  }
  if (ROBUST) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cost_a += __shfl_down_sync(0xffffffffu, cost_a, o);
      cost_f += __shfl_down_sync(0xffffffffu, cost_f, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sred[2 * w] = cost_a; sred[2 * w + 1] = cost_f; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double sa = 0, sf = 0;
      for (int k = 0; k < kEvalThreads / 32; ++k) { sa += sred[2 * k]; sf += sred[2 * k + 1]; }
      a.cost_part[2 * blockIdx.x] = sa;
      a.cost_part[2 * blockIdx.x + 1] = sf;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
PointArgs make_point_args(const tslam_dev_problem* d, bool unweighted) {
  PointArgs a{};
  a.cams = d->cams.p; a.rho = d->rho.p;
  a.uv = reinterpret_cast<const double2*>(d->p_uv.p); a.ray = reinterpret_cast<const double2*>(d->p_ray.p);
  a.cam = d->p_cam.p; a.host = d->p_host.p; a.lm = d->p_lm.p;
  a.fx = d->K_point[0]; a.fy = d->K_point[1]; a.cx = d->K_point[2]; a.cy = d->K_point[3];
  a.wx = unweighted ? 1.0 : d->w_point[0]; a.wy = unweighted ? 1.0 : d->w_point[1];
  a.n = d->n_pobs;
  a.huber = 0; a.active = nullptr; a.cost_part = nullptr;
  return a;
}
TextArgs make_text_args(const tslam_dev_problem* d, bool unweighted) {
  TextArgs a{};
  a.cams = d->cams.p; a.theta = d->theta.p;
  a.rays = reinterpret_cast<const double2*>(d->t_rays.p); a.iref = d->t_iref.p;
  a.musigma = reinterpret_cast<const double2*>(d->t_musigma.p);
  a.cam = d->t_cam.p; a.host = d->t_host.p; a.plane = d->t_plane.p; a.img = d->t_img.p;
  a.imgs = d->imgs.p; a.img_w = d->img_w; a.img_h = d->img_h;
  a.fx = d->K_text[0]; a.fy = d->K_text[1]; a.cx = d->K_text[2]; a.cy = d->K_text[3];
  a.wT = unweighted ? 1.0 : d->w_text;
  a.n = d->n_tobs; a.free_mask = 7u; a.free_masks = nullptr;
  a.huber = 0; a.active = nullptr; a.cost_part = nullptr;
  return a;
}

int launch_eval_points(tslam_ctx* ctx, tslam_dev_problem* d, int kind, bool want_J) {
  if (d->n_pobs == 0) return TSLAM_OK;
  const int ncols = kind == TSLAM_PT_POSE ? 6 : (kind == TSLAM_PT_RHO ? 1 : 13);
  TSL_CUDA(d->pr.reserve(2 * (size_t)d->n_pobs));
  if (want_J) { TSL_CUDA(d->pJ.reserve((size_t)d->n_pobs * 2 * ncols)); d->pJ_cols = ncols; }
  PointArgs a = make_point_args(d, kind == TSLAM_PT_BA_NW || kind == TSLAM_PT_RHO);
  const int grid = (d->n_pobs + kEvalThreads - 1) / kEvalThreads;
  double2* r = reinterpret_cast<double2*>(d->pr.p);
  if (!want_J) LAUNCH(launch_k(point_eval_kernel<0, 13, false, false>, grid, kEvalThreads, 0, ctx->stream, a, r, nullptr));
  else if (ncols == 13) {
    // occupancy variant: 8 CTAs/SM (64 registers, ~48 B of L1-resident spills) vs 6 CTAs/SM (79 registers)
    static const bool occ8 = getenv("TSLAM_EVAL_OCC8") != nullptr;   // measured: 8 CTAs/SM is slower (0.67 vs 0.81 of HBM peak, profiles/r1_notes.md)
    if (occ8) LAUNCH(launch_k(point_eval_kernel<0, 13, true, false, 8>, grid, kEvalThreads, 0, ctx->stream, a, r, d->pJ.p));
    else LAUNCH(launch_k(point_eval_kernel<0, 13, true, false, 6>, grid, kEvalThreads, 0, ctx->stream, a, r, d->pJ.p));
  }
  else if (ncols == 6) LAUNCH(launch_k(point_eval_kernel<0, 6, true, false>, grid, kEvalThreads, 0, ctx->stream, a, r, d->pJ.p));
  else LAUNCH(launch_k(point_eval_kernel<12, 1, true, false>, grid, kEvalThreads, 0, ctx->stream, a, r, d->pJ.p));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

template <int MODE>
static void launch_text_mode(tslam_ctx* ctx, tslam_dev_problem* d, int kind, const TextArgs& a, int grid) {
  if (kind == TSLAM_TX_BA) LAUNCH(launch_k(text_eval_kernel<0, 15, MODE, true, false>, grid, kEvalThreads, 0, ctx->stream, a, d->tr.p, d->tJ.p));
  else if (kind == TSLAM_TX_POSE) LAUNCH(launch_k(text_eval_kernel<0, 6, MODE, true, false>, grid, kEvalThreads, 0, ctx->stream, a, d->tr.p, d->tJ.p));
  else LAUNCH(launch_k(text_eval_kernel<12, 3, MODE, true, false>, grid, kEvalThreads, 0, ctx->stream, a, d->tr.p, d->tJ.p));
}

int launch_eval_text(tslam_ctx* ctx, tslam_dev_problem* d, int kind, int jac_mode, bool want_J) {
  if (d->n_tobs == 0) return TSLAM_OK;
  if (want_J && jac_mode == TSLAM_JAC_ANALYTIC_TMA) return launch_eval_text_tma(ctx, d, kind);
  const int ncols = kind == TSLAM_TX_BA ? 15 : (kind == TSLAM_TX_POSE ? 6 : 3);
  TSL_CUDA(d->tr.reserve(8 * (size_t)d->n_tobs));
  if (want_J) { TSL_CUDA(d->tJ.reserve((size_t)d->n_tobs * 8 * ncols)); d->tJ_cols = ncols; }
  TextArgs a = make_text_args(d, kind == TSLAM_TX_THETA);
  a.free_mask = kind == TSLAM_TX_BA ? 7u : (kind == TSLAM_TX_POSE ? 1u : 4u);
  const int grid = (8 * d->n_tobs + kEvalThreads - 1) / kEvalThreads;
  if (!want_J) LAUNCH(launch_k(text_eval_kernel<0, 15, 0, false, false>, grid, kEvalThreads, 0, ctx->stream, a, d->tr.p, nullptr));
  else if (jac_mode == TSLAM_JAC_CENTRAL_DIFF) launch_text_mode<TSLAM_JAC_CENTRAL_DIFF>(ctx, d, kind, a, grid);
  else launch_text_mode<TSLAM_JAC_ANALYTIC>(ctx, d, kind, a, grid);
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

// solver-facing launchers (robustified, 13 / 15 columns) -- declared in solver.cuh
int launch_eval_points_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* rho, const uint8_t* active,
                              double* r, double* J, double* cost_part, int* n_parts) {
  *n_parts = 0;
  if (d->n_pobs == 0) return TSLAM_OK;
  PointArgs a = make_point_args(d, false);
  a.cams = cams; a.rho = rho; a.huber = d->huber_point; a.active = active; a.cost_part = cost_part;
  const int grid = (d->n_pobs + kEvalThreads - 1) / kEvalThreads;
  *n_parts = grid;
  if (J) LAUNCH(launch_k(point_eval_kernel<0, 13, true, true>, grid, kEvalThreads, 0, ctx->stream, a, reinterpret_cast<double2*>(r), J));
  else LAUNCH(launch_k(point_eval_kernel<0, 13, false, true>, grid, kEvalThreads, 0, ctx->stream, a, reinterpret_cast<double2*>(r), nullptr));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}
int launch_eval_text_robust(tslam_ctx* ctx, tslam_dev_problem* d, const double* cams, const double* theta, const uint8_t* active,
                            const uint8_t* free_masks, int jac_mode, double* r, double* J, double* cost_part, int* n_parts) {
  *n_parts = 0;
  if (d->n_tobs == 0) return TSLAM_OK;
  TextArgs a = make_text_args(d, false);
  a.cams = cams; a.theta = theta; a.huber = d->huber_text; a.active = active; a.cost_part = cost_part; a.free_masks = free_masks;
  const int grid = (8 * d->n_tobs + kEvalThreads - 1) / kEvalThreads;
  *n_parts = grid;
  if (!J) LAUNCH(launch_k(text_eval_kernel<0, 15, 0, false, true>, grid, kEvalThreads, 0, ctx->stream, a, r, nullptr));
  else if (jac_mode == TSLAM_JAC_CENTRAL_DIFF) LAUNCH(launch_k(text_eval_kernel<0, 15, TSLAM_JAC_CENTRAL_DIFF, true, true>, grid, kEvalThreads, 0, ctx->stream, a, r, J));
  else LAUNCH(launch_k(text_eval_kernel<0, 15, TSLAM_JAC_ANALYTIC, true, true>, grid, kEvalThreads, 0, ctx->stream, a, r, J));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl
