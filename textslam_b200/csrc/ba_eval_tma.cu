// TMA-staged variant of the text residual + Jacobian kernel (nume_BAText / nume_PoseOptimText / nume_thetaText in the analytic
// Jacobian mode; include/nume_BAText.h:58-91): BASELINE.json's north star names "TMA-staged image patches and camera/landmark
// parameter blocks in shared memory" for this path, so it exists beside the __ldg-tap kernel of ba_eval.cu and both are measured
// (bench.py text_on, profiles/r2_notes.md) — the faster one is what tslam_solve uses.
//
// One CTA per run of consecutive text blocks that share (observing camera, host camera, plane, image) — the 25 features of
// one text object seen in one keyframe in the reference's insertion order (src/optimizer.cc:1447-1510): the two camera blocks
// and the plane live in shared memory, every lane projects its pattern pixel, the CTA reduces the bounding box of the 2x2 tap
// footprints, ONE cp.async.bulk.tensor.2d (TMA) brings that window of the u8 image into shared memory (zero fill outside the
// image, completion on an mbarrier) and the taps are read from there. Windows larger than the TMA box fall back to global taps.
// Results are identical to text_eval_kernel's (same arithmetic on the same tap values).
#include <cuda.h>
#include "ctx.cuh"
#include "solver.cuh"
#include "ba_device.cuh"

namespace tsl {

constexpr int TMA_BOX_W = 128, TMA_BOX_H = 64;   // bytes x rows of the staged window (u8 image): 8 KB
constexpr int TMA_THREADS = 256;                 // up to 32 text blocks per run

struct TextTmaArgs {
  const double* cams; const double* theta;
  const double2* rays; const double* iref; const double2* musigma;
  const int32_t* cam; const int32_t* host; const int32_t* plane; const int32_t* img;
  const uint8_t* imgs; int img_w, img_h;
  double fx, fy, cx, cy, wT;
  const int32_t* run_ptr; int n_runs;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int COL0, int NCOLS>
__global__ void __launch_bounds__(TMA_THREADS) text_eval_tma_kernel(const __grid_constant__ CUtensorMap tmap, TextTmaArgs a, double* __restrict__ r_out,
                                                                    double* __restrict__ J_out) {
  PDL_PROLOGUE();
  __shared__ __align__(128) uint8_t tile[TMA_BOX_W * TMA_BOX_H];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ double s_par[17];                  // [cam 7 | host 7 | theta 3] of the run
  __shared__ int s_box[4];                      // min x, min y, max x, max y of the tap footprints
  const int run = blockIdx.x, tid = threadIdx.x;
  const int b0 = a.run_ptr[run], nb = a.run_ptr[run + 1] - b0;
  if (tid == 0) {
    s_box[0] = s_box[1] = 0x7fffffff; s_box[2] = s_box[3] = -0x7fffffff;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < 7) s_par[tid] = __ldg(a.cams + 7 * (size_t)__ldg(a.cam + b0) + tid);
  else if (tid < 14) s_par[tid] = __ldg(a.cams + 7 * (size_t)__ldg(a.host + b0) + tid - 7);
  else if (tid < 17) s_par[tid] = __ldg(a.theta + 3 * (size_t)__ldg(a.plane + b0) + tid - 14);
  __syncthreads();
  const int lb = tid >> 3, px = tid & 7;
  const bool valid = lb < nb;
  const int b = b0 + lb;
  const size_t gpx = (size_t)b * 8 + px;
  Cam c, h;
#pragma unroll
  for (int k = 0; k < 4; ++k) { c.q[k] = s_par[k]; h.q[k] = s_par[7 + k]; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { c.t[k] = s_par[4 + k]; h.t[k] = s_par[11 + k]; }
  const double th[3] = {s_par[14], s_par[15], s_par[16]};
  double2 ray = make_double2(0.0, 0.0), ms = make_double2(0.0, 1.0);
  double iref = 0.0;
  if (valid) { ray = a.rays[gpx]; iref = a.iref[gpx]; ms = a.musigma[b]; }
  // ---- footprint of this lane's taps ----
  RelPose P;
  relative_pose(c.q, c.t, h.q, h.t, P);
  {
    const double rho = -(ray.x * th[0] + ray.y * th[1] + th[2]);
    const double X = (P.R[0] * ray.x + P.R[1] * ray.y + P.R[2]) / rho + P.t[0];
    const double Y = (P.R[3] * ray.x + P.R[4] * ray.y + P.R[5]) / rho + P.t[1];
    const double Z = (P.R[6] * ray.x + P.R[7] * ray.y + P.R[8]) / rho + P.t[2];
    const double u = a.fx * X / Z + a.cx, v = a.fy * Y / Z + a.cy;
    if (valid && u >= 0.0 && v >= 0.0 && u < (double)a.img_w && v < (double)a.img_h) {   // pixels outside the image read nothing
      const int uf = (int)floor(u), vf = (int)floor(v);
      // one pixel of margin on every side: the functor re-derives (u, v) and may round to the neighbouring cell
      atomicMin(&s_box[0], uf - 1); atomicMin(&s_box[1], vf - 1); atomicMax(&s_box[2], uf + 2); atomicMax(&s_box[3], vf + 2);
    }
  }
  __syncthreads();
  const int x0 = s_box[0] & ~15, y0 = s_box[1];   // 16-byte aligned window start: full 128-bit shared-memory writes by the copy engine
  const bool staged = s_box[2] >= s_box[0] && s_box[2] - x0 < TMA_BOX_W && s_box[3] - y0 < TMA_BOX_H;
  const int img = __ldg(a.img + b0);
  if (staged && tid == 0) {
    const unsigned bar = smem_u32(&mbar), dst = smem_u32(tile);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TMA_BOX_W * TMA_BOX_H) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(bar), "r"(x0), "r"(img * a.img_h + y0) : "memory");
  }
  if (staged) {
    unsigned ok = 0;
    const unsigned bar = smem_u32(&mbar);
    do {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    } while (!ok);
  }
  TextImg im{a.imgs + (size_t)img * a.img_w * a.img_h, a.img_w, a.img_h};
  if (staged) { im.tile = tile; im.tx0 = x0; im.ty0 = y0; im.tw = TMA_BOX_W; }
  double res = 0.0, Jr[15];
  if (valid) res = text_pixel_analytic(c, h, th, ray.x, ray.y, im, a.fx, a.fy, a.cx, a.cy, ms.x, ms.y, iref, a.wT, Jr);
  if (valid) {
    r_out[gpx] = res;
    double* out = J_out + gpx * NCOLS;
#pragma unroll
    for (int k = 0; k < NCOLS; ++k) out[k] = Jr[COL0 + k];
  }
}

// One tensor map per device problem: the u8 image stack as a 2-D tensor (width, n_imgs * height).
static int make_tensor_map(const tslam_dev_problem* d, CUtensorMap* tm) {
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static PFN encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TSL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return set_error(TSLAM_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<PFN>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)d->img_w, (cuuint64_t)d->n_imgs * (cuuint64_t)d->img_h};
  const cuuint64_t gstride[1] = {(cuuint64_t)d->img_w};
  const cuuint32_t box[2] = {TMA_BOX_W, TMA_BOX_H};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d->imgs.p, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TSLAM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TSLAM_OK;
}

bool text_tma_supported(const tslam_dev_problem* d) { return d->n_truns > 0 && d->img_w % 16 == 0 && d->n_imgs > 0; }

int launch_eval_text_tma(tslam_ctx* ctx, tslam_dev_problem* d, int kind) {
  if (d->n_tobs == 0) return TSLAM_OK;
  if (!text_tma_supported(d)) return set_error(TSLAM_ERR_ARG, "TMA text path needs the run table of a device-resident problem and an image width that is a multiple of 16");
  const int ncols = kind == TSLAM_TX_BA ? 15 : (kind == TSLAM_TX_POSE ? 6 : 3);
  TSL_CUDA(d->tr.reserve(8 * (size_t)d->n_tobs));
  TSL_CUDA(d->tJ.reserve((size_t)d->n_tobs * 8 * ncols)); d->tJ_cols = ncols;
  CUtensorMap tm;
  int rc = make_tensor_map(d, &tm);
  if (rc) return rc;
  TextTmaArgs a{};
  a.cams = d->cams.p; a.theta = d->theta.p;
  a.rays = reinterpret_cast<const double2*>(d->t_rays.p); a.iref = d->t_iref.p; a.musigma = reinterpret_cast<const double2*>(d->t_musigma.p);
  a.cam = d->t_cam.p; a.host = d->t_host.p; a.plane = d->t_plane.p; a.img = d->t_img.p;
  a.imgs = d->imgs.p; a.img_w = d->img_w; a.img_h = d->img_h;
  a.fx = d->K_text[0]; a.fy = d->K_text[1]; a.cx = d->K_text[2]; a.cy = d->K_text[3];
  a.wT = kind == TSLAM_TX_THETA ? 1.0 : d->w_text;
  a.run_ptr = d->t_run_ptr.p; a.n_runs = d->n_truns;
  if (kind == TSLAM_TX_BA) LAUNCH(launch_k(text_eval_tma_kernel<0, 15>, d->n_truns, TMA_THREADS, 0, ctx->stream, tm, a, d->tr.p, d->tJ.p));
  else if (kind == TSLAM_TX_POSE) LAUNCH(launch_k(text_eval_tma_kernel<0, 6>, d->n_truns, TMA_THREADS, 0, ctx->stream, tm, a, d->tr.p, d->tJ.p));
  else LAUNCH(launch_k(text_eval_tma_kernel<12, 3>, d->n_truns, TMA_THREADS, 0, ctx->stream, tm, a, d->tr.p, d->tJ.p));
  TSL_CHECK_LAUNCH();
  return TSLAM_OK;
}

}  // namespace tsl
