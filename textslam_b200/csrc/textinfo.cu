// mu / sigma of the projected text quad on the GPU: replaces tool::CalTextinfo + CalStatistics
// (/root/reference/src/tool.cc:1178-1262), which the optimizer calls once per (text observation x pyramid level) at
// every problem build (src/optimizer.cc:1182-1184,1490-1492,1525-1527,1952-1954,2185-2188) with a full-image
// cv::Mat::zeros + cv::fillPoly. One CTA per quad: the fillPoly mask (8-connected outline via cv::clipLine +
// cv::LineIterator, 16.16 fixed-point scan-line interior) is built as a bitmap in shared memory, then the
// intensities under it are reduced (integer sum -> mu exactly as the reference; two-pass variance).
// Same rasteriser as oracle/textinfo_oracle.cpp (pinned against cv2 there; see its header for the border caveat).
#include "ctx.cuh"

namespace tsl {
typedef long long i64;

__device__ bool ti_clip_line(int w, int h, i64& x1, i64& y1, i64& x2, i64& y2) {
  const i64 right = w - 1, bottom = h - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    i64 a;
    if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (i64)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1)); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
    if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (i64)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1)); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) { a = c1 == 1 ? 0 : right; y1 += (i64)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1)); x1 = a; c1 = 0; }
      if (c2) { a = c2 == 1 ? 0 : right; y2 += (i64)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1)); x2 = a; c2 = 0; }
    }
  }
  return (c1 | c2) == 0;
}

struct TiEdge { int y0, y1; i64 x, dx; };
struct TiLine { i64 x1, y1, x2, y2; int visible; };

__global__ void __launch_bounds__(256) text_info_kernel(const uint8_t* __restrict__ imgs, int w, int h, const double* __restrict__ quads,
                                                        const int* __restrict__ quad_img, int nq, double* __restrict__ mu_out,
                                                        double* __restrict__ sigma_out, int* __restrict__ ok_out) {
  extern __shared__ unsigned bitmap[];          // h rows x wpr words
  __shared__ TiEdge edges[4];
  __shared__ TiLine lines[4];
  __shared__ int s_ne, bbox[4];
  __shared__ long long s_sum[256];
  __shared__ int s_cnt[256];
  __shared__ double s_dbl[256];
  const int q = blockIdx.x;
  if (q >= nq) return;
  const int wpr = (w + 31) >> 5;
  const uint8_t* img = imgs + (size_t)quad_img[q] * w * h;
  const double* Q = quads + 8 * (size_t)q;
  if (threadIdx.x == 0) {
    int px[4], py[4];
    int xMin = w + 1, xMax = -1, yMin = h + 1, yMax = -1;
    for (int i = 0; i < 4; ++i) {
      const double vx = Q[2 * i], vy = Q[2 * i + 1];
      px[i] = (int)vx; py[i] = (int)vy;
      if (vx > xMax) xMax = (int)ceil(vx);
      if (vx < xMin) xMin = (int)floor(vx);
      if (vy > yMax) yMax = (int)ceil(vy);
      if (vy < yMin) yMin = (int)floor(vy);
    }
    if (xMin < 0) xMin = 0; if (xMin >= w) xMin = w - 1;
    if (yMin < 0) yMin = 0; if (yMin >= h) yMin = h - 1;
    if (xMax >= w) xMax = w - 1; if (xMax < 0) xMax = 0;
    if (yMax >= h) yMax = h - 1; if (yMax < 0) yMax = 0;
    bbox[0] = xMin; bbox[1] = xMax; bbox[2] = yMin; bbox[3] = yMax;
    int ne = 0;
    for (int i = 0; i < 4; ++i) {
      const int j = (i + 3) & 3;
      const i64 ax = px[j], ay = py[j], bx = px[i], by = py[i];
      i64 cx1 = ax, cy1 = ay, cx2 = bx, cy2 = by;
      const bool vis = ti_clip_line(w, h, cx1, cy1, cx2, cy2);
      lines[i].x1 = cx1; lines[i].y1 = cy1; lines[i].x2 = cx2; lines[i].y2 = cy2; lines[i].visible = vis ? 1 : 0;
      if (ay == by) continue;
      const bool outside = ax < 0 || ax >= w || bx < 0 || bx >= w || ay < 0 || ay >= h || by < 0 || by >= h;
      i64 e0x = ax << 16, e0y = ay, e1x = bx << 16, e1y = by;
      if (outside && cy1 != cy2) { e0x = cx1 << 16; e0y = cy1; e1x = cx2 << 16; e1y = cy2; }
      TiEdge e;
      e.dx = (e1x - e0x) / (e1y - e0y);
      if (ay < by) { e.y0 = (int)ay; e.y1 = (int)by; e.x = e0x + (ay - e0y) * e.dx; }
      else { e.y0 = (int)by; e.y1 = (int)ay; e.x = e1x + (by - e1y) * e.dx; }
      edges[ne++] = e;
    }
    s_ne = ne;
  }
  for (int e = threadIdx.x; e < wpr * h; e += 256) bitmap[e] = 0u;
  __syncthreads();
  // ---- scan-line interior: one thread per row ----
  const int ne = s_ne;
  if (ne > 0) {
    int ymin = edges[0].y0, ymax = edges[0].y1;
    for (int k = 1; k < ne; ++k) { ymin = min(ymin, edges[k].y0); ymax = max(ymax, edges[k].y1); }
    ymax = min(ymax, h);
    for (int y = max(ymin, 0) + threadIdx.x; y < ymax; y += 256) {
      i64 xs[4]; int n = 0;
      for (int k = 0; k < ne; ++k) if (edges[k].y0 <= y && y < edges[k].y1) xs[n++] = edges[k].x + (i64)(y - edges[k].y0) * edges[k].dx;
      for (int a = 1; a < n; ++a) { const i64 v = xs[a]; int b = a - 1; while (b >= 0 && xs[b] > v) { xs[b + 1] = xs[b]; --b; } xs[b + 1] = v; }
      for (int k = 0; k + 1 < n; k += 2) {
        i64 x1 = (xs[k] + 65535) >> 16, x2 = xs[k + 1] >> 16;
        if (x1 < w && x2 >= 0) {
          if (x1 < 0) x1 = 0;
          if (x2 > w - 1) x2 = w - 1;
          for (int x = (int)x1; x <= (int)x2; ++x) bitmap[y * wpr + (x >> 5)] |= 1u << (x & 31);
        }
      }
    }
  }
  __syncthreads();
  // ---- outline: cv::Line = LineIterator(8-connected, leftToRight) on the clipped segment, one thread per edge ----
  if (threadIdx.x < 4 && lines[threadIdx.x].visible) {
    i64 x1 = lines[threadIdx.x].x1, y1 = lines[threadIdx.x].y1, x2 = lines[threadIdx.x].x2, y2 = lines[threadIdx.x].y2;
    i64 dx = x2 - x1, dy = y2 - y1;
    int by = 1;
    if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }
    if (dy < 0) { dy = -dy; by = -1; }
    const bool swp = dy > dx;
    if (swp) { const i64 t = dx; dx = dy; dy = t; }
    i64 err = dx - (dy + dy); const i64 plus = dx + dx, minus = -(dy + dy);
    i64 x = x1, y = y1;
    for (i64 i = 0; i <= dx; ++i) {
      atomicOr(&bitmap[(int)y * wpr + ((int)x >> 5)], 1u << ((int)x & 31));
      const bool m = err < 0;
      err += minus + (m ? plus : 0);
      if (swp) { y += by; if (m) x += 1; } else { x += 1; if (m) y += by; }
    }
  }
  __syncthreads();
  // ---- statistics over bbox ∩ mask ----
  const int xMin = bbox[0], xMax = bbox[1], yMin = bbox[2], yMax = bbox[3];
  const int bw = xMax - xMin + 1, bh = yMax - yMin + 1;
  long long sum = 0; int cnt = 0;
  for (int e = threadIdx.x; e < bw * bh; e += 256) {
    const int y = yMin + e / bw, x = xMin + e % bw;
    if (bitmap[y * wpr + (x >> 5)] & (1u << (x & 31))) { sum += img[(size_t)y * w + x]; ++cnt; }
  }
  s_sum[threadIdx.x] = sum; s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; } __syncthreads(); }
  const int n = s_cnt[0];
  if (n == 0) { if (threadIdx.x == 0) { mu_out[q] = 0.0; sigma_out[q] = 0.0; ok_out[q] = 0; } return; }
  const double mu = (double)s_sum[0] / (double)n;
  double ss = 0.0;
  for (int e = threadIdx.x; e < bw * bh; e += 256) {
    const int y = yMin + e / bw, x = xMin + e % bw;
    if (bitmap[y * wpr + (x >> 5)] & (1u << (x & 31))) { const double d = (double)img[(size_t)y * w + x] - mu; ss += d * d; }
  }
  s_dbl[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s_dbl[threadIdx.x] += s_dbl[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) {
    const double sg = sqrt(s_dbl[0] / (double)(n - 1));
    mu_out[q] = mu; sigma_out[q] = sg; ok_out[q] = (sg != 0.0) ? 1 : 0;   // n == 1 -> 0/0 = NaN -> "!= 0" is true in the reference too
  }
}
}  // namespace tsl

using namespace tsl;

extern "C" int tslam_text_info(tslam_ctx* ctx, const uint8_t* imgs, int n_imgs, int w, int h, const double* quads, const int32_t* quad_img, int n_quads,
                               double* mu_out, double* sigma_out, int32_t* ok_out) {
  if (!ctx || !imgs || !quads || !quad_img || !mu_out || !sigma_out || !ok_out) return set_error(TSLAM_ERR_ARG, "null argument");
  if (n_quads <= 0) return TSLAM_OK;
  if (w < 1 || h < 1 || n_imgs < 1) return set_error(TSLAM_ERR_ARG, "bad image size");
  for (int i = 0; i < n_quads; ++i) if ((unsigned)quad_img[i] >= (unsigned)n_imgs) return set_error(TSLAM_ERR_ARG, "quad %d: image index out of range", i);
  const size_t smem = (size_t)((w + 31) / 32) * h * sizeof(unsigned);
  if (smem > 200 * 1024) return set_error(TSLAM_ERR_ARG, "image %dx%d too large for the shared-memory mask (%zu B)", w, h, smem);
  TSL_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (smem > ctx->attr_textinfo_smem) { TSL_CUDA(cudaFuncSetAttribute(text_info_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); ctx->attr_textinfo_smem = smem; }
  DevBuf<uint8_t> dimg; DevBuf<double> dq, dmu, dsg; DevBuf<int> dqi, dok;
  TSL_CUDA(dimg.upload(imgs, (size_t)n_imgs * w * h, st)); TSL_CUDA(dq.upload(quads, 8 * (size_t)n_quads, st)); TSL_CUDA(dqi.upload(quad_img, n_quads, st));
  TSL_CUDA(dmu.reserve(n_quads)); TSL_CUDA(dsg.reserve(n_quads)); TSL_CUDA(dok.reserve(n_quads));
  LAUNCH(text_info_kernel<<<n_quads, 256, smem, st>>>(dimg.p, w, h, dq.p, dqi.p, n_quads, dmu.p, dsg.p, dok.p));
  TSL_CHECK_LAUNCH();
  TSL_CUDA(cudaMemcpyAsync(mu_out, dmu.p, sizeof(double) * n_quads, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(sigma_out, dsg.p, sizeof(double) * n_quads, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(ok_out, dok.p, sizeof(int) * n_quads, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaStreamSynchronize(st));
  return TSLAM_OK;
}
