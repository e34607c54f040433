// Post-solve chi^2 gates: the outlier loops after Problem::Evaluate in PyrPoseOptim
// (/root/reference/src/optimizer.cc:1236-1302) and PyrBA (:1616-1684), evaluated on the final residuals while they
// are in HBM. Integer/flag outputs: every comparison is the reference's own double expression ((r/w)*(r/w) > chi2,
// |r/w| > chi2, (double)bad/(double)size > ratio) with IEEE division, so the flags are bit-identical to the host loops.
#include <vector>
#include "ctx.cuh"
#include "solver.cuh"

namespace tsl {

__global__ void gate_points_kernel(int n, const double* __restrict__ r, double wx, double wy, double chi2, uint8_t* __restrict__ bad,
                                   int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool b = false;
  if (i < n) {
    const double2 v = reinterpret_cast<const double2*>(r)[i];
    const double qx = v.x / wx, qy = v.y / wy;
    const double chix = __dmul_rn(qx, qx), chiy = __dmul_rn(qy, qy);
    b = chix > chi2 || chiy > chi2;
    bad[i] = b ? 1 : 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counts[0], __popc(m));
}

// 8 lanes per text block (one pixel residual each)
__global__ void gate_text_kernel(int n, const double* __restrict__ r, double wt, double chi2, const int* __restrict__ t_obj, int n_obj,
                                 uint8_t* __restrict__ bad, int* __restrict__ obj_bad_blocks, int* __restrict__ obj_blocks,
                                 int* __restrict__ counts, int* __restrict__ err) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int blk = (int)(g >> 3);
  bool over = false;
  if (blk < n) over = fabs(r[g] / wt) > chi2;
  const unsigned m = __ballot_sync(0xffffffffu, over);
  const int sub = (threadIdx.x & 31) >> 3;
  const bool b = ((m >> (8 * sub)) & 0xffu) != 0;
  if (blk < n && (threadIdx.x & 7) == 0) {
    bad[blk] = b ? 1 : 0;
    const int o = t_obj[blk];
    if (o < 0 || o >= n_obj) { atomicExch(err, 1); return; }
    atomicAdd(&obj_blocks[o], 1);
    if (b) { atomicAdd(&obj_bad_blocks[o], 1); atomicAdd(&counts[1], 1); }
  }
}

__global__ void gate_objects_kernel(int n_obj, const int* __restrict__ obj_size, const int* __restrict__ obj_bad_blocks,
                                    const int* __restrict__ obj_blocks, double ratio, uint8_t* __restrict__ obj_bad,
                                    int* __restrict__ counts, int* __restrict__ err) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_obj) return;
  const int sz = obj_size[o];
  if (obj_blocks[o] != sz) { atomicExch(err, 2); obj_bad[o] = 0; return; }
  bool b = false;
  if (sz > 0) b = ((double)obj_bad_blocks[o] / (double)sz) > ratio;
  obj_bad[o] = b ? 1 : 0;
  if (b) atomicAdd(&counts[2], 1);
}

// d_rp: 2*n_pobs, d_rt: 8*n_tobs device residuals. Host outputs may be NULL.
int gate_device(tslam_ctx* ctx, const double* d_rp, const double* d_rt, int n_pobs, int n_tobs, const int32_t* t_obj,
                const int32_t* obj_size, int n_obj, const tslam_gate_options* g, uint8_t* pt_bad, uint8_t* tf_bad, uint8_t* obj_bad,
                int32_t* counts_out) {
  if (!g) return set_error(TSLAM_ERR_ARG, "gate options missing");
  if (n_pobs < 0 || n_tobs < 0 || n_obj < 0) return set_error(TSLAM_ERR_ARG, "negative size");
  const bool do_p = g->gate_points && n_pobs > 0, do_t = g->gate_text && n_tobs > 0;
  if (do_p && !pt_bad) return set_error(TSLAM_ERR_ARG, "pt_bad missing");
  if (do_t && (!tf_bad || !t_obj || (n_obj > 0 && (!obj_size || !obj_bad)))) return set_error(TSLAM_ERR_ARG, "text gate arrays missing");
  if (do_p && (g->w_point[0] == 0.0 || g->w_point[1] == 0.0)) return set_error(TSLAM_ERR_ARG, "w_point == 0");
  if (do_t && g->w_text == 0.0) return set_error(TSLAM_ERR_ARG, "w_text == 0");
  cudaStream_t st = ctx->stream;
  DevBuf<uint8_t> dpb, dtb, dob; DevBuf<int> dto, dsz, dcnt;
  // dcnt: [0..2] counts, [3] error flag, then n_obj bad-block counters and n_obj block counters
  const size_t ncnt = 4 + 2 * (size_t)n_obj;
  TSL_CUDA(dcnt.reserve(ncnt));
  TSL_CUDA(cudaMemsetAsync(dcnt.p, 0, ncnt * sizeof(int), st));
  if (do_p) {
    double chi2 = g->chi2_mono;
    if (g->relax_below_text_blocks > 0 && n_tobs < g->relax_below_text_blocks) chi2 = g->chi2_mono + g->relax_amount;
    TSL_CUDA(dpb.reserve(n_pobs));
    LAUNCH(gate_points_kernel<<<(n_pobs + 255) / 256, 256, 0, st>>>(n_pobs, d_rp, g->w_point[0], g->w_point[1], chi2, dpb.p, dcnt.p));
    TSL_CHECK_LAUNCH();
    TSL_CUDA(cudaMemcpyAsync(pt_bad, dpb.p, n_pobs, cudaMemcpyDeviceToHost, st));
  }
  if (do_t) {
    TSL_CUDA(dtb.reserve(n_tobs));
    TSL_CUDA(dto.upload(t_obj, n_tobs, st));
    const long long threads = 8ll * n_tobs;
    LAUNCH(gate_text_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(n_tobs, d_rt, g->w_text, g->chi2_text, dto.p, n_obj, dtb.p,
                                                                             dcnt.p + 4, dcnt.p + 4 + n_obj, dcnt.p, dcnt.p + 3));
    TSL_CHECK_LAUNCH();
    TSL_CUDA(cudaMemcpyAsync(tf_bad, dtb.p, n_tobs, cudaMemcpyDeviceToHost, st));
    if (n_obj > 0) {
      TSL_CUDA(dob.reserve(n_obj));
      TSL_CUDA(dsz.upload(obj_size, n_obj, st));
      LAUNCH(gate_objects_kernel<<<(n_obj + 127) / 128, 128, 0, st>>>(n_obj, dsz.p, dcnt.p + 4, dcnt.p + 4 + n_obj, g->text_ratio, dob.p, dcnt.p,
                                                                    dcnt.p + 3));
      TSL_CHECK_LAUNCH();
      TSL_CUDA(cudaMemcpyAsync(obj_bad, dob.p, n_obj, cudaMemcpyDeviceToHost, st));
    }
  }
  int h[4] = {0, 0, 0, 0};
  TSL_CUDA(cudaMemcpyAsync(h, dcnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  if (h[3] == 1) return set_error(TSLAM_ERR_ARG, "t_obj entry outside [0, n_obj)");
  if (h[3] == 2) return set_error(TSLAM_ERR_ARG, "obj_size[o] differs from the number of text blocks naming object o");
  if (counts_out) { counts_out[0] = h[0]; counts_out[1] = h[1]; counts_out[2] = h[2]; }
  return TSLAM_OK;
}

}  // namespace tsl

using namespace tsl;

extern "C" int tslam_gate_residuals(tslam_ctx* ctx, const double* final_residuals, int n_pobs, int n_tobs, const int32_t* t_obj,
                                    const int32_t* obj_size, int n_obj, const tslam_gate_options* gate, uint8_t* pt_bad, uint8_t* tf_bad,
                                    uint8_t* obj_bad, int32_t counts_out[3]) {
  if (!ctx || !gate) return set_error(TSLAM_ERR_ARG, "null argument");
  if (n_pobs < 0 || n_tobs < 0 || n_obj < 0) return set_error(TSLAM_ERR_ARG, "negative size");
  if (counts_out) counts_out[0] = counts_out[1] = counts_out[2] = 0;
  const size_t total = 2 * (size_t)n_pobs + 8 * (size_t)n_tobs;
  if (total == 0) return TSLAM_OK;
  if (!final_residuals) return set_error(TSLAM_ERR_ARG, "final_residuals missing");
  TSL_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> r;
  TSL_CUDA(r.upload(final_residuals, total, ctx->stream));
  return gate_device(ctx, r.p, r.p + 2 * (size_t)n_pobs, n_pobs, n_tobs, t_obj, obj_size, n_obj, gate, pt_bad, tf_bad, obj_bad, counts_out);
}
