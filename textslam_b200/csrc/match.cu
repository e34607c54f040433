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
}  // namespace tsl

using namespace tsl;

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
