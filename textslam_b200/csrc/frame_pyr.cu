// Direct-method frame pyramid on the GPU: replaces frame::GetPyrMat
// (/root/reference/src/frame.cc:178-202) — cv::pyrDown chain (factor 2.0, `iScaleLevels` levels), cv::Sobel into
// CV_8U in x and y (negative derivatives saturate to 0, as in the reference: ddepth = img.type()), and
// cv::addWeighted(grad_x, 0.5, grad_y, 0.5). Bit-exact against the oracle, which is pinned against cv2.
// Layout: per image one tight record holding, per level, four u8 planes [img | grad | grad_x | grad_y].
#include <vector>
#include "ctx.cuh"

namespace tsl {
__device__ __forceinline__ int fp_reflect101(int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; return p; }

__global__ void pyrdown_kernel(const uint8_t* src, uint8_t* dst, size_t rec_bytes, size_t src_off, size_t dst_off,
                               int sw, int sh, int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw) return;
  const uint8_t* s = src + (size_t)blockIdx.z * rec_bytes + src_off;
  int xi[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) xi[i] = fp_reflect101(2 * x + i - 2, sw);
  int acc = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint8_t* row = s + (size_t)fp_reflect101(2 * y + j - 2, sh) * sw;
    const int rs = row[xi[0]] + 4 * row[xi[1]] + 6 * row[xi[2]] + 4 * row[xi[3]] + row[xi[4]];
    acc += (j == 0 || j == 4) ? rs : ((j == 2) ? 6 * rs : 4 * rs);
  }
  dst[(size_t)blockIdx.z * rec_bytes + dst_off + (size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
}

__global__ void sobel_kernel(uint8_t* __restrict__ rec, size_t rec_bytes, size_t off, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  uint8_t* base = rec + (size_t)blockIdx.z * rec_bytes + off;
  const size_t plane = (size_t)w * h;
  const uint8_t* r0 = base + (size_t)fp_reflect101(y - 1, h) * w; const uint8_t* r1 = base + (size_t)y * w; const uint8_t* r2 = base + (size_t)fp_reflect101(y + 1, h) * w;
  const int xm = fp_reflect101(x - 1, w), xp = fp_reflect101(x + 1, w);
  const int dx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
  const int dy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
  const int a = min(255, max(0, dx)), b = min(255, max(0, dy));
  const size_t o = (size_t)y * w + x;
  base[2 * plane + o] = (uint8_t)a;
  base[3 * plane + o] = (uint8_t)b;
  base[plane + o] = (uint8_t)min(255, __double2int_rn((double)(a + b) * 0.5));   // addWeighted: round half to even
}
}  // namespace tsl

struct tslam_frame_pyr {
  tslam_ctx* ctx = nullptr;
  int nlevels = 0, w = 0, h = 0, n_alloc = 0, last_n = 0;
  std::vector<int> lw, lh;
  std::vector<size_t> off;
  size_t rec_bytes = 0;
  tsl::DevBuf<uint8_t> rec;
};

using namespace tsl;

extern "C" {
int tslam_frame_pyr_create(tslam_ctx* ctx, int nlevels, tslam_frame_pyr** out) {
  if (!ctx || !out || nlevels < 1 || nlevels > 16) return set_error(TSLAM_ERR_ARG, "bad argument");
  tslam_frame_pyr* p = new tslam_frame_pyr();
  p->ctx = ctx; p->nlevels = nlevels;
  *out = p;
  return TSLAM_OK;
}
void tslam_frame_pyr_destroy(tslam_frame_pyr* p) { if (p) { cudaSetDevice(p->ctx->device); delete p; } }

int tslam_frame_pyr_build(tslam_frame_pyr* p, const uint8_t* const* imgs, int n_imgs, int w, int hgt, int stride) {
  if (!p || !imgs || n_imgs <= 0 || stride < w || w < 2 || hgt < 2) return set_error(TSLAM_ERR_ARG, "bad argument");
  TSL_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t st = p->ctx->stream;
  if (p->w != w || p->h != hgt) {
    p->w = w; p->h = hgt; p->n_alloc = 0;
    p->lw.assign(p->nlevels, 0); p->lh.assign(p->nlevels, 0); p->off.assign(p->nlevels, 0);
    size_t off = 0; int cw = w, ch = hgt;
    for (int l = 0; l < p->nlevels; ++l) {
      if (cw < 2 || ch < 2) return set_error(TSLAM_ERR_ARG, "image too small for %d levels", p->nlevels);
      p->lw[l] = cw; p->lh[l] = ch; p->off[l] = off;
      off += 4 * (size_t)cw * ch; off = (off + 15) & ~(size_t)15;
      cw = (cw + 1) / 2; ch = (ch + 1) / 2;
    }
    p->rec_bytes = off;
  }
  if (n_imgs > p->n_alloc) { TSL_CUDA(p->rec.reserve((size_t)n_imgs * p->rec_bytes)); p->n_alloc = n_imgs; }
  for (int i = 0; i < n_imgs; ++i)
    TSL_CUDA(cudaMemcpy2DAsync(p->rec.p + (size_t)i * p->rec_bytes, w, imgs[i], stride, w, hgt, cudaMemcpyHostToDevice, st));
  for (int l = 1; l < p->nlevels; ++l)
    LAUNCH(pyrdown_kernel<<<dim3((p->lw[l] + 127) / 128, p->lh[l], n_imgs), 128, 0, st>>>(p->rec.p, p->rec.p, p->rec_bytes, p->off[l - 1], p->off[l],
                                                                                           p->lw[l - 1], p->lh[l - 1], p->lw[l], p->lh[l]));
  for (int l = 0; l < p->nlevels; ++l)
    LAUNCH(sobel_kernel<<<dim3((p->lw[l] + 127) / 128, p->lh[l], n_imgs), 128, 0, st>>>(p->rec.p, p->rec_bytes, p->off[l], p->lw[l], p->lh[l]));
  TSL_CHECK_LAUNCH();
  TSL_CUDA(cudaStreamSynchronize(st));
  p->last_n = n_imgs;
  return TSLAM_OK;
}

int tslam_frame_pyr_level_size(tslam_frame_pyr* p, int level, int* w, int* hgt) {
  if (!p || level < 0 || level >= p->nlevels || p->lw.empty()) return set_error(TSLAM_ERR_ARG, "no pyramid yet / bad level");
  *w = p->lw[level]; *hgt = p->lh[level];
  return TSLAM_OK;
}

int tslam_frame_pyr_get(tslam_frame_pyr* p, int img, int level, int what, uint8_t* out) {
  if (!p || !out || level < 0 || level >= p->nlevels || img < 0 || img >= p->last_n || what < 0 || what > 3) return set_error(TSLAM_ERR_ARG, "bad argument");
  TSL_CUDA(cudaSetDevice(p->ctx->device));
  const size_t plane = (size_t)p->lw[level] * p->lh[level];
  TSL_CUDA(cudaMemcpyAsync(out, p->rec.p + (size_t)img * p->rec_bytes + p->off[level] + (size_t)what * plane, plane, cudaMemcpyDeviceToHost, p->ctx->stream));
  TSL_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return TSLAM_OK;
}
}  // extern "C"
