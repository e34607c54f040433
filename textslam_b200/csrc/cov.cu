// Plane-parameter covariance: replaces the ceres::Covariance block of PyrThetaOptim
// (/root/reference/src/optimizer.cc:2219-2238): (J_theta' J_theta)^-1 of the 3x3 theta block at the solution,
// Jacobian loss-corrected like Ceres' Covariance (apply_loss_function default). Valid when theta blocks do not
// couple with other free blocks, which is how the reference uses it (only theta is free in PyrThetaOptim).
#include <vector>
#include "ctx.cuh"
#include "solver.cuh"

namespace tsl {
__global__ void theta_cov_kernel(int n_planes, const int* __restrict__ ptr, const int* __restrict__ blocks, const double* __restrict__ J /*n x 8 x 15*/,
                                 double* __restrict__ cov, int* __restrict__ singular) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_planes) return;
  double V[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = ptr[p]; e < ptr[p + 1]; ++e) {
    const double* Jb = J + (size_t)blocks[e] * 120;
    for (int row = 0; row < 8; ++row) {
      const double a = Jb[row * 15 + 12], b = Jb[row * 15 + 13], c = Jb[row * 15 + 14];
      V[0] += a * a; V[1] += a * b; V[2] += a * c; V[4] += b * b; V[5] += b * c; V[8] += c * c;
    }
  }
  V[3] = V[1]; V[6] = V[2]; V[7] = V[5];
  const double c00 = V[4] * V[8] - V[5] * V[5], c01 = V[2] * V[5] - V[1] * V[8], c02 = V[1] * V[5] - V[2] * V[4];
  const double det = V[0] * c00 + V[1] * c01 + V[2] * c02;
  double* o = cov + 9 * (size_t)p;
  if (!(det > 0.0) || !(V[0] > 0.0)) { for (int k = 0; k < 9; ++k) o[k] = 0.0; atomicAdd(singular, 1); return; }
  const double id = 1.0 / det;
  o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
  o[3] = o[1]; o[4] = (V[0] * V[8] - V[2] * V[2]) * id; o[5] = (V[1] * V[2] - V[0] * V[5]) * id;
  o[6] = o[2]; o[7] = o[5]; o[8] = (V[0] * V[4] - V[1] * V[1]) * id;
}
}  // namespace tsl

using namespace tsl;

extern "C" int tslam_theta_covariance(tslam_ctx* ctx, const tslam_ba_problem* p, int jac_mode, double* cov_out, int32_t* n_singular) {
  if (!ctx || !p || !cov_out) return set_error(TSLAM_ERR_ARG, "null argument");
  if (p->n_planes <= 0) return TSLAM_OK;
  TSL_CUDA(cudaSetDevice(ctx->device));
  tslam_dev_problem d;
  int rc = upload_problem(ctx, p, &d, false);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  DevBuf<double> r, J, parts, cov; DevBuf<int> dptr, dblk, dsing;
  TSL_CUDA(r.reserve(8 * (size_t)d.n_tobs)); TSL_CUDA(J.reserve(120 * (size_t)d.n_tobs));
  TSL_CUDA(parts.reserve(2 * ((8 * (size_t)d.n_tobs + 127) / 128) + 2)); TSL_CUDA(cov.reserve(9 * (size_t)d.n_planes)); TSL_CUDA(dsing.reserve(1));
  int nparts = 0;
  if ((rc = launch_eval_text_robust(ctx, &d, d.cams.p, d.theta.p, nullptr, nullptr, jac_mode, r.p, J.p, parts.p, &nparts))) return rc;
  // text blocks per plane (CSR, insertion order)
  std::vector<int> ptr(p->n_planes + 1, 0), blk(p->n_tobs);
  for (int i = 0; i < p->n_tobs; ++i) ptr[p->t_plane[i] + 1]++;
  for (int k = 0; k < p->n_planes; ++k) ptr[k + 1] += ptr[k];
  { std::vector<int> cur(ptr.begin(), ptr.end() - 1); for (int i = 0; i < p->n_tobs; ++i) blk[cur[p->t_plane[i]]++] = i; }
  TSL_CUDA(dptr.upload(ptr.data(), ptr.size(), st)); TSL_CUDA(dblk.upload(blk.data(), blk.size(), st));
  TSL_CUDA(cudaMemsetAsync(dsing.p, 0, sizeof(int), st));
  LAUNCH(theta_cov_kernel<<<(d.n_planes + 127) / 128, 128, 0, st>>>(d.n_planes, dptr.p, dblk.p, J.p, cov.p, dsing.p));
  TSL_CHECK_LAUNCH();
  int ns = 0;
  TSL_CUDA(cudaMemcpyAsync(cov_out, cov.p, sizeof(double) * 9 * (size_t)d.n_planes, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(&ns, dsing.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  if (n_singular) *n_singular = ns;
  return TSLAM_OK;
}
