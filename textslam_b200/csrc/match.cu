// Projection-guided descriptor matching core: the inner loop of tracking::SearchFrom3D /
// SearchFrom3DAdd / SearchFrom3DLocalTrack (/root/reference/src/tracking.cc:1161-1175, 1241-1256, 1310-1325) —
// for every query descriptor, the FIRST candidate (in list order) with the minimum Hamming distance
// (tracking::DescriptorDistance, :2762-2778: 256-bit popcount). Candidate lists (GetFeaturesInArea) and the
// sequential uniqueness bookkeeping stay on the host. Integer work: bit-exact.
#include "ctx.cuh"

namespace tsl {
__global__ void __launch_bounds__(128) match_hamming_kernel(const uint32_t* __restrict__ q, const uint32_t* __restrict__ t, const int* __restrict__ cand_ptr,
                                                            const int* __restrict__ cand, int nq, int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                            int* __restrict__ second_dist) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nq) return;
  uint32_t qd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) qd[k] = __ldg(q + 8 * (size_t)w + k);
  // key = dist << 20 | position in the list  -> min key == smallest distance, earliest candidate on ties
  unsigned best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;
  const int e0 = cand_ptr[w], e1 = cand_ptr[w + 1];
  for (int e = e0 + lane; e < e1; e += 32) {
    const uint32_t* td = t + 8 * (size_t)cand[e];
    int d = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) d += __popc(qd[k] ^ __ldg(td + k));
    const unsigned key = ((unsigned)d << 20) | (unsigned)min(e - e0, 0xFFFFF);
    if (key < best) { second = best; best = key; } else if (key < second) second = key;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
    const unsigned nb = min(best, ob);
    const unsigned ns = min(max(best, ob), min(second, os));
    best = nb; second = ns;
  }
  if (lane == 0) {
    if (best == 0xFFFFFFFFu) { best_idx[w] = -1; best_dist[w] = 2147483647; second_dist[w] = 2147483647; }
    else {
      best_idx[w] = cand[e0 + (int)(best & 0xFFFFFu)];
      best_dist[w] = (int)(best >> 20);
      second_dist[w] = second == 0xFFFFFFFFu ? 2147483647 : (int)(second >> 20);
    }
  }
}

// ---- tracking::SearchFrom3D* up to the uniqueness bookkeeping (src/tracking.cc:1124-1176, 1206-1256, 1282-1327): project the map
// point into the frame, frame::GetFeaturesInArea (src/frame.cc:415-468) over the 64 x 48 grid, first minimum-distance candidate.
// One warp per map point: the projection is evaluated by every lane (FP64, same expression order as the reference:
// K (R_cr ray / rho + t_cr), u = x / z), the query window in FP32 exactly as GetFeaturesInArea computes it (the double u, v are
// narrowed to float at the call, src/tracking.cc:1154), cells in the reference's order (ix outer, iy inner, insertion order inside
// a cell) so that "first candidate wins a tie" holds: key = distance << 20 | position in that order.
struct Search3dArgs {
  const double* Tcw; double fx, fy, cx, cy;
  int n_pts; const double2* ray; const double* rho; const double* poses; const int* host; const int* query;
  const uint32_t* qdesc; const float2* kp_xy; const int* kp_oct; const uint32_t* tdesc;
  int cols, rows; float min_x, min_y, max_x, max_y, inv_w, inv_h; const int* cell_ptr; const int* cell_idx;
  float radius; int min_level, max_level;
  int* best_idx; int* best_dist; double2* uv;
  // SearchFrom3DLocalTrack (src/tracking.cc:1282-1345): projections given (mapPts::LocalTrackProj), key points already matched to a
  // well-observed map point are skipped (:1311-1313), the runner-up distance is wanted for the ratio test (:1331-1334)
  const double2* uv_in; const uint8_t* kp_skip; int* second_dist;
};
__device__ __forceinline__ void quat_R(const double* q, double R[9]) {   // unit quaternion (w, x, y, z) -> rotation, as ba_device.cuh
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
__global__ void __launch_bounds__(128) search3d_kernel(Search3dArgs a) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= a.n_pts) return;
  const int qi = a.query[w];
  int out_idx = -1, out_dist = 2147483647, out_second = 2147483647;
  double u = 0.0, v = 0.0;
  bool live = qi >= 0;
  if (live && a.uv_in) { u = a.uv_in[w].x; v = a.uv_in[w].y; }
  else if (live) {
    double Rc[9], Rr[9];
    quat_R(a.Tcw, Rc);
    const double* Pr = a.poses + 7 * (size_t)a.host[w];
    quat_R(Pr, Rr);
    // T_cr = T_cw T_rw^-1: R_cr = R_c R_r', t_cr = t_c - R_cr t_r
    double R[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[3 * i + j] = __dadd_rn(__dadd_rn(__dmul_rn(Rc[3 * i], Rr[3 * j]), __dmul_rn(Rc[3 * i + 1], Rr[3 * j + 1])), __dmul_rn(Rc[3 * i + 2], Rr[3 * j + 2]));
    const double2 ry = a.ray[w];
    const double ir = 1.0 / a.rho[w];
    double p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double Rt = __dadd_rn(__dadd_rn(__dmul_rn(R[3 * i], Pr[4]), __dmul_rn(R[3 * i + 1], Pr[5])), __dmul_rn(R[3 * i + 2], Pr[6]));
      const double Rray = __dadd_rn(__dadd_rn(__dmul_rn(R[3 * i], ry.x), __dmul_rn(R[3 * i + 1], ry.y)), R[3 * i + 2]);
      p[i] = __dadd_rn(__dmul_rn(ir, Rray), __dadd_rn(a.Tcw[4 + i], -Rt));
    }
    const double X = __dadd_rn(__dmul_rn(a.fx, p[0]), __dmul_rn(a.cx, p[2])), Y = __dadd_rn(__dmul_rn(a.fy, p[1]), __dmul_rn(a.cy, p[2]));
    u = X / p[2]; v = Y / p[2];
    if (u < (double)a.min_x || u > (double)a.max_x || v < (double)a.min_y || v > (double)a.max_y) live = false;
  }
  if (live) {
    const float x = (float)u, y = (float)v, r = a.radius;
    const int c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, a.min_x), r), a.inv_w)));
    const int c1 = min(a.cols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, a.min_x), r), a.inv_w)));
    const int r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, a.min_y), r), a.inv_h)));
    const int r1 = min(a.rows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, a.min_y), r), a.inv_h)));
    if (c0 < a.cols && c1 >= 0 && r0 < a.rows && r1 >= 0) {
      uint32_t qd[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) qd[k] = __ldg(a.qdesc + 8 * (size_t)qi + k);
      const bool check = a.min_level > 0 || a.max_level >= 0;
      unsigned best = 0xFFFFFFFFu, second = 0xFFFFFFFFu; int best_k = -1, base = 0;
      for (int ix = c0; ix <= c1; ++ix)
        for (int iy = r0; iy <= r1; ++iy) {
          const int c = ix * a.rows + iy;
          const int e0 = a.cell_ptr[c], e1 = a.cell_ptr[c + 1];
          for (int e = e0 + lane; e < e1; e += 32) {
            const int k = a.cell_idx[e];
            if (check) {
              const int oc = a.kp_oct[k];
              if (oc < a.min_level) continue;
              if (a.max_level >= 0 && oc > a.max_level) continue;
            }
            const float2 kp = a.kp_xy[k];
            if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
            if (a.kp_skip && a.kp_skip[k]) continue;
            const uint32_t* td = a.tdesc + 8 * (size_t)k;
            int d = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) d += __popc(qd[j] ^ __ldg(td + j));
            const unsigned key = ((unsigned)d << 20) | (unsigned)min(base + e - e0, 0xFFFFF);
            if (key < best) { second = best; best = key; best_k = k; } else if (key < second) second = key;
          }
          base += e1 - e0;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o); const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        second = min(max(best, ob), min(second, os));   // second smallest of the union: the reference's bestDist2 is the multiset's runner-up
        if (ob < best) { best = ob; best_k = ok; }
      }
      if (best != 0xFFFFFFFFu) { out_idx = best_k; out_dist = (int)(best >> 20); }
      if (second != 0xFFFFFFFFu) out_second = (int)(second >> 20);
    }
  }
  if (lane == 0) {
    a.best_idx[w] = out_idx; a.best_dist[w] = out_dist;
    if (a.second_dist) a.second_dist[w] = out_second;
    if (a.uv) a.uv[w] = make_double2(u, v);
  }
}
}  // namespace tsl

using namespace tsl;

static int search_impl(tslam_ctx* ctx, const double* Tcw, const double* K, int n_pts, const double* pt_ray, const double* pt_rho,
                       const double* poses, int n_poses, const int32_t* pt_host, const int32_t* pt_query, const uint8_t* query_desc, int n_query,
                       const float* kp_xy, const int32_t* kp_octave, const uint8_t* train_desc, int n_kp, const tslam_frame_grid* g,
                       float radius, int min_level, int max_level, int32_t* best_idx, int32_t* best_dist, double* uv_out,
                       const double* uv_in, const uint8_t* kp_skip, int32_t* second_dist) {
  if (!ctx || !g || !best_idx || !best_dist) return set_error(TSLAM_ERR_ARG, "null argument");
  if (n_pts <= 0) return TSLAM_OK;
  if (!pt_query || !query_desc || !g->cell_ptr) return set_error(TSLAM_ERR_ARG, "null argument");
  if (!uv_in && (!Tcw || !K || !pt_ray || !pt_rho || !poses || !pt_host)) return set_error(TSLAM_ERR_ARG, "null argument");
  if (n_kp > 0 && (!kp_xy || !kp_octave || !train_desc || !g->cell_idx)) return set_error(TSLAM_ERR_ARG, "null keypoint array");
  if (g->cols <= 0 || g->rows <= 0 || g->cols * (long long)g->rows > (1 << 20)) return set_error(TSLAM_ERR_ARG, "bad grid dimensions");
  const int ncell = g->cols * g->rows;
  if (g->cell_ptr[0] != 0) return set_error(TSLAM_ERR_ARG, "cell_ptr[0] must be 0");
  for (int c = 0; c < ncell; ++c) if (g->cell_ptr[c + 1] < g->cell_ptr[c]) return set_error(TSLAM_ERR_ARG, "cell_ptr not monotone at cell %d", c);
  const int nent = g->cell_ptr[ncell];
  for (int e = 0; e < nent; ++e) if ((unsigned)g->cell_idx[e] >= (unsigned)n_kp) return set_error(TSLAM_ERR_ARG, "grid entry %d out of range", e);
  for (int i = 0; i < n_pts; ++i) {
    if (!uv_in && (unsigned)pt_host[i] >= (unsigned)n_poses) return set_error(TSLAM_ERR_ARG, "map point %d: host pose out of range", i);
    if (pt_query[i] >= n_query) return set_error(TSLAM_ERR_ARG, "map point %d: query descriptor out of range", i);
  }
  TSL_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  AllocStreamScope alloc_scope(st);
  DevBuf<double> dT, dray, drho, dposes, duv, duvin; DevBuf<int> dhost, dq, doct, dcp, dci, dbi, dbd, dsd; DevBuf<uint8_t> dqd, dtd, dskip; DevBuf<float> dxy;
  if (uv_in) TSL_CUDA(duvin.upload(uv_in, 2 * (size_t)n_pts, st));
  else {
    TSL_CUDA(dT.upload(Tcw, 7, st)); TSL_CUDA(dray.upload(pt_ray, 2 * (size_t)n_pts, st)); TSL_CUDA(drho.upload(pt_rho, n_pts, st));
    TSL_CUDA(dposes.upload(poses, 7 * (size_t)n_poses, st)); TSL_CUDA(dhost.upload(pt_host, n_pts, st));
  }
  if (kp_skip && n_kp > 0) TSL_CUDA(dskip.upload(kp_skip, n_kp, st));
  TSL_CUDA(dsd.reserve(n_pts));
  TSL_CUDA(dq.upload(pt_query, n_pts, st));
  TSL_CUDA(dqd.upload(query_desc, 32 * (size_t)n_query, st)); TSL_CUDA(dxy.upload(kp_xy, 2 * (size_t)n_kp, st)); TSL_CUDA(doct.upload(kp_octave, n_kp, st));
  TSL_CUDA(dtd.upload(train_desc, 32 * (size_t)n_kp, st)); TSL_CUDA(dcp.upload(g->cell_ptr, (size_t)ncell + 1, st)); TSL_CUDA(dci.upload(g->cell_idx, nent, st));
  TSL_CUDA(dbi.reserve(n_pts)); TSL_CUDA(dbd.reserve(n_pts)); TSL_CUDA(duv.reserve(2 * (size_t)n_pts));
  Search3dArgs a;
  a.Tcw = dT.p; a.fx = K ? K[0] : 0; a.fy = K ? K[1] : 0; a.cx = K ? K[2] : 0; a.cy = K ? K[3] : 0;
  a.uv_in = uv_in ? reinterpret_cast<const double2*>(duvin.p) : nullptr; a.kp_skip = (kp_skip && n_kp > 0) ? dskip.p : nullptr; a.second_dist = dsd.p;
  a.n_pts = n_pts; a.ray = reinterpret_cast<const double2*>(dray.p); a.rho = drho.p; a.poses = dposes.p; a.host = dhost.p; a.query = dq.p;
  a.qdesc = reinterpret_cast<const uint32_t*>(dqd.p); a.kp_xy = reinterpret_cast<const float2*>(dxy.p); a.kp_oct = doct.p; a.tdesc = reinterpret_cast<const uint32_t*>(dtd.p);
  a.cols = g->cols; a.rows = g->rows; a.min_x = g->min_x; a.min_y = g->min_y; a.max_x = g->max_x; a.max_y = g->max_y; a.inv_w = g->inv_w; a.inv_h = g->inv_h;
  a.cell_ptr = dcp.p; a.cell_idx = dci.p; a.radius = radius; a.min_level = min_level; a.max_level = max_level;
  a.best_idx = dbi.p; a.best_dist = dbd.p; a.uv = uv_out ? reinterpret_cast<double2*>(duv.p) : nullptr;
  LAUNCH(search3d_kernel<<<(n_pts * 32 + 127) / 128, 128, 0, st>>>(a));
  TSL_CHECK_LAUNCH();
  TSL_CUDA(cudaMemcpyAsync(best_idx, dbi.p, sizeof(int) * n_pts, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(best_dist, dbd.p, sizeof(int) * n_pts, cudaMemcpyDeviceToHost, st));
  if (uv_out) TSL_CUDA(cudaMemcpyAsync(uv_out, duv.p, sizeof(double) * 2 * n_pts, cudaMemcpyDeviceToHost, st));
  if (second_dist) TSL_CUDA(cudaMemcpyAsync(second_dist, dsd.p, sizeof(int) * n_pts, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  return TSLAM_OK;
}

extern "C" int tslam_match_hamming(tslam_ctx* ctx, const uint8_t* query_desc, int n_query, const uint8_t* train_desc, int n_train,
                                   const int32_t* cand_ptr, const int32_t* cand_idx, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist) {
  if (!ctx || !query_desc || !train_desc || !cand_ptr || !best_idx || !best_dist) return set_error(TSLAM_ERR_ARG, "null argument");
  if (n_query <= 0) return TSLAM_OK;
  if (cand_ptr[0] != 0) return set_error(TSLAM_ERR_ARG, "cand_ptr[0] must be 0");
  const int nc = cand_ptr[n_query];
  for (int i = 0; i < n_query; ++i)
    if (cand_ptr[i + 1] < cand_ptr[i] || cand_ptr[i + 1] - cand_ptr[i] > 0xFFFFF) return set_error(TSLAM_ERR_ARG, "bad candidate list of query %d", i);
  for (int e = 0; e < nc; ++e) if ((unsigned)cand_idx[e] >= (unsigned)n_train) return set_error(TSLAM_ERR_ARG, "candidate %d out of range", e);
  TSL_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevBuf<uint8_t> dq, dt; DevBuf<int> dp, dc, dbi, dbd, dsd;
  TSL_CUDA(dq.upload(query_desc, 32 * (size_t)n_query, st)); TSL_CUDA(dt.upload(train_desc, 32 * (size_t)n_train, st));
  TSL_CUDA(dp.upload(cand_ptr, (size_t)n_query + 1, st)); TSL_CUDA(dc.upload(cand_idx, nc, st));
  TSL_CUDA(dbi.reserve(n_query)); TSL_CUDA(dbd.reserve(n_query)); TSL_CUDA(dsd.reserve(n_query));
  LAUNCH(match_hamming_kernel<<<(n_query * 32 + 127) / 128, 128, 0, st>>>(reinterpret_cast<const uint32_t*>(dq.p), reinterpret_cast<const uint32_t*>(dt.p), dp.p, dc.p,
                                                                         n_query, dbi.p, dbd.p, dsd.p));
  TSL_CHECK_LAUNCH();
  TSL_CUDA(cudaMemcpyAsync(best_idx, dbi.p, sizeof(int) * n_query, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(best_dist, dbd.p, sizeof(int) * n_query, cudaMemcpyDeviceToHost, st));
  if (second_dist) TSL_CUDA(cudaMemcpyAsync(second_dist, dsd.p, sizeof(int) * n_query, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  return TSLAM_OK;
}

extern "C" int tslam_search_from_3d(tslam_ctx* ctx, const double* Tcw, const double* K, int n_pts, const double* pt_ray, const double* pt_rho,
                                    const double* poses, int n_poses, const int32_t* pt_host, const int32_t* pt_query, const uint8_t* query_desc, int n_query,
                                    const float* kp_xy, const int32_t* kp_octave, const uint8_t* train_desc, int n_kp, const tslam_frame_grid* g,
                                    float radius, int min_level, int max_level, int32_t* best_idx, int32_t* best_dist, double* uv_out) {
  if (!Tcw || !K) return set_error(TSLAM_ERR_ARG, "null argument");
  return search_impl(ctx, Tcw, K, n_pts, pt_ray, pt_rho, poses, n_poses, pt_host, pt_query, query_desc, n_query, kp_xy, kp_octave, train_desc, n_kp, g, radius,
                     min_level, max_level, best_idx, best_dist, uv_out, nullptr, nullptr, nullptr);
}

extern "C" int tslam_search_in_area(tslam_ctx* ctx, int n_pts, const double* uv, const int32_t* pt_query, const uint8_t* query_desc, int n_query,
                                    const float* kp_xy, const int32_t* kp_octave, const uint8_t* kp_skip, const uint8_t* train_desc, int n_kp,
                                    const tslam_frame_grid* g, float radius, int min_level, int max_level, int32_t* best_idx, int32_t* best_dist,
                                    int32_t* second_dist) {
  if (n_pts > 0 && !uv) return set_error(TSLAM_ERR_ARG, "null argument");
  return search_impl(ctx, nullptr, nullptr, n_pts, nullptr, nullptr, nullptr, 0, nullptr, pt_query, query_desc, n_query, kp_xy, kp_octave, train_desc, n_kp, g, radius,
                     min_level, max_level, best_idx, best_dist, nullptr, uv, kp_skip, second_dist);
}
