// ORB extractor (FAST + quad-tree + IC angle + rBRIEF) for batches of images on sm_100a.
//
// Replaces TextSLAM::ORBextractor::operator() and everything below it
// (/root/reference/src/ORBextractor.cc:1054-1116: ComputePyramid :1118-1143, ComputeKeyPointsOctTree
// :766-854, DistributeOctTree :540-764, IC_Angle :77-104, GaussianBlur + computeOrbDescriptor
// :1096-1101,:108-147) and the OpenCV primitives called there (cv::resize, cv::FAST, cv::GaussianBlur,
// cv::fastAtan2, cvRound), bit-exactly (checked against oracle/orb_oracle.cpp).
//
// Data layout in HBM (all u8 planes are tight, [image][row][col]):
//   pyr[l], blur[l], score[l]   per level, n_imgs * h_l * w_l bytes each
//   slots[img][cell][256]       packed candidates (x_rel:10 | y_rel:10 | response:8), row-major inside a cell,
//   cell_count[img][cell]       so that (cell order, slot order) == vToDistributeKeys order of the reference
//   sel[img][level][cap]        quad-tree winners in list order, then keypoints / descriptors level-major
// The 19-px replicated border of mvImagePyramid (:1124-1139) is never read by any consumer (FAST runs on
// cell ROIs inside the 16-px margin, IC_Angle / BRIEF stay >= 4 px inside) and is therefore not stored.
//
// Kernels: resize (one thread per output pixel, OpenCV's 11-bit fixed point), fast_score (FAST-9/16
// corner measure per pixel), cell_nms (one CTA per 30-px cell: threshold fallback 20 -> 7, 3x3 non-max
// suppression, ordered compaction), distribute (one CTA per (image, level): the quad-tree), blur7
// (separable fixed-point Gaussian), orient_describe (one warp per keypoint).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cfloat>
#include "ctx.cuh"
#include "orb_math.h"

namespace tsl {

constexpr int ORB_EDGE = 19, ORB_HALF_PATCH = 15, ORB_PATCH = 31;
constexpr int CELL_CAP = 256;       // candidates per cell after NMS (<= ceil(31*32/4))
constexpr int MAX_LEVELS = 16;
constexpr int MAXID = 2048;         // quad-tree node ids per (image, level)
constexpr int MAXCELLS = 512;       // cells per level

__constant__ int c_umax[ORB_HALF_PATCH + 1];
// the rBRIEF pattern as one 32-bit word (x0, y0, x1, y1) per (comparison k, lane) in GLOBAL memory: every lane reads its own entries — a
// coalesced, L1-resident load here, 32 serialised fetches per access from the constant cache (orient_describe_kernel: 509 -> 260 us per batch)
__device__ uint32_t g_pat_packed[8 * 32];
static const int8_t h_pattern[1024] = {
#include "orb_pattern.inc"
};

static inline int cv_round_host(double v) { return (int)std::nearbyint(v); }

struct CellRect { short x0, y0, x1, y1, offx, offy; };  // ROI [x0,x1) x [y0,y1) in level coords; offx = j*wCell, offy = i*hCell

struct LevelInfo {
  int w, h;
  int minBX, minBY, maxBX, maxBY;
  int ncells, cell_base;     // cells of this level; index of its first cell among all levels
  int nfeat;                 // mnFeaturesPerLevel
  float scale;               // mvScaleFactor
  int patch;                 // (int)(31 * scale)
  size_t plane_off;          // byte offset of this level inside one image's pyramid record
};

}  // namespace tsl

struct tslam_orb {
  tslam_ctx* ctx = nullptr;
  int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0, blur_variant = 0;
  float scaleFactor = 1.2f;
  std::vector<float> mvScale, mvInvScale;
  std::vector<int> perLevel;
  int umax[tsl::ORB_HALF_PATCH + 1];
  // per input size
  int w = 0, h = 0, n_alloc = 0;
  std::vector<tsl::LevelInfo> L;
  int total_cells = 0, sel_cap = 0, out_cap = 0;
  size_t img_bytes = 0;   // bytes of one image's whole pyramid record
  tsl::DevBuf<uint8_t> pyr, blur, score;
  tsl::DevBuf<uint32_t> slots;
  tsl::DevBuf<int> cell_count, sel_count, err;
  tsl::DevBuf<uint16_t> node_of;
  tsl::DevBuf<uint8_t> kq;
  tsl::DevBuf<uint32_t> sel;
  tsl::DevBuf<tsl::CellRect> cells;
  tsl::DevBuf<tsl::LevelInfo> Ld;
  tsl::DevBuf<int> xofs, yofs;      // resize tables, concatenated per level
  tsl::DevBuf<short> xa, ya;        // (a0,a1) interleaved
  std::vector<int> xtab_off, ytab_off;
  tsl::DevBuf<tslam_keypoint> kp;
  tsl::DevBuf<uint8_t> desc;
  tsl::DevBuf<int> counts;
  int last_n = 0;
  // second stream: the 7x7 Gaussian of every level (needed by the descriptors only) runs beside the quad-tree distribution, whose one CTA
  // of 8 warps per SM (115 KB of shared memory each) leaves the SMs mostly idle
  cudaStream_t s2 = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_level[16] = {}, ev_res[16] = {};
  ~tslam_orb() {
    if (s2) cudaStreamDestroy(s2);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (cudaEvent_t e : ev_level) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_res) if (e) cudaEventDestroy(e);
  }
};

namespace tsl {

// ---------------------------------------------------------------------------------------------------
// resize: cv::resize(..., INTER_LINEAR) for 8UC1
// ---------------------------------------------------------------------------------------------------
__global__ void resize_kernel(const uint8_t* pyr, uint8_t* out_base, size_t img_bytes, size_t src_off, size_t dst_off,
                              int sw, int sh, int dw, int dh, const int* __restrict__ xofs, const short* __restrict__ xa,
                              const int* __restrict__ yofs, const short* __restrict__ ya) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dw) return;
  const uint8_t* src = pyr + (size_t)blockIdx.z * img_bytes + src_off;
  uint8_t* dst = out_base + (size_t)blockIdx.z * img_bytes + dst_off;
  const int sx0 = xofs[x], sx1 = min(sx0 + 1, sw - 1), sy0 = yofs[y], sy1 = min(sy0 + 1, sh - 1);
  const int a0 = xa[2 * x], a1 = xa[2 * x + 1], b0 = ya[2 * y], b1 = ya[2 * y + 1];
  const uint8_t* p0 = src + (size_t)sy0 * sw; const uint8_t* p1 = src + (size_t)sy1 * sw;
  const int r0 = a0 * p0[sx0] + a1 * p0[sx1];
  const int r1 = a0 * p1[sx0] + a1 * p1[sx1];
  dst[(size_t)y * dw + x] = (uint8_t)((((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2);
}

// ---------------------------------------------------------------------------------------------------
// FAST-9/16 corner measure m(p) = max over the 16 arcs of 9 of min(c_k - p), min(p - c_k); stored as
// u8 (0 when m <= min_th: such pixels are corners at neither threshold). Corner at t <=> m > t,
// cv::FAST response = m - 1.
// ---------------------------------------------------------------------------------------------------
// The measure is evaluated through 16-bit arc masks: the largest t for which 9 contiguous ring pixels are all
// > v + t (or all < v - t) is found by bisection on t; a mask has a 9-run iff m & rot(m,1) & ... & rot(m,8) != 0.
__device__ __forceinline__ bool has_run9(unsigned m) {
  m |= m << 16;                       // unroll the ring
  unsigned r = m & (m >> 1);          // runs of 2
  r &= r >> 2;                        // runs of 4
  r &= r >> 4;                        // runs of 8
  r &= m >> 8;                        // runs of 9
  return (r & 0xFFFFu) != 0u;
}
__device__ __forceinline__ int fast_measure_bisect(const uint8_t* p, int stride, int min_th) {
  const int v = p[0];
  int ring[16];
  ring[0] = p[3 * stride]; ring[8] = p[-3 * stride]; ring[4] = p[3]; ring[12] = p[-3];
  // every arc of 9 contains pixel 0 or 8, and pixel 4 or 12: most pixels are rejected after 5 loads
  if ((abs(ring[0] - v) <= min_th && abs(ring[8] - v) <= min_th) || (abs(ring[4] - v) <= min_th && abs(ring[12] - v) <= min_th)) return 0;
  ring[1] = p[3 * stride + 1]; ring[2] = p[2 * stride + 2]; ring[3] = p[stride + 3];
  ring[5] = p[-stride + 3]; ring[6] = p[-2 * stride + 2]; ring[7] = p[-3 * stride + 1];
  ring[9] = p[-3 * stride - 1]; ring[10] = p[-2 * stride - 2]; ring[11] = p[-stride - 3];
  ring[13] = p[stride - 3]; ring[14] = p[2 * stride - 2]; ring[15] = p[3 * stride - 1];
  // corner at threshold t  <=>  m > t ; find the largest t in [min_th, 254] that is still a corner -> m = t + 1
  auto corner = [&](int t) {
    unsigned mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { mb |= (ring[k] > v + t ? 1u : 0u) << k; md |= (ring[k] < v - t ? 1u : 0u) << k; }
    return has_run9(mb) || has_run9(md);
  };
  if (!corner(min_th)) return 0;
  int lo = min_th, hi = 255;            // corner(lo) true, corner(hi) false (|diff| <= 255)
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (corner(mid)) lo = mid; else hi = mid; }
  return lo + 1;
}

// The same measure in closed form: m = max(m_bright, m_dark), m_bright = max over the 16 arcs of the minimum of (c_k - p) over the
// arc's 9 pixels (m_dark with p - c_k). The 16 sliding-window minima over the circular ring come from a doubling ladder
// (windows of 2, 4, 8, then 8 + 1): 4 x 16 integer min per polarity instead of 8 bisection steps x 32 comparisons.
// corner at t <=> m > t, so the bisection's result (largest corner threshold + 1) is m itself, 0 when m <= min_th.
__device__ __forceinline__ int ring_max_of_window9_min(const int (&d)[16]) {
  int w2[16], w4[16], w8[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
  for (int k = 0; k < 16; ++k) w4[k] = min(w2[k], w2[(k + 2) & 15]);
#pragma unroll
  for (int k = 0; k < 16; ++k) w8[k] = min(w4[k], w4[(k + 4) & 15]);
  int best = -256;
#pragma unroll
  for (int k = 0; k < 16; ++k) best = max(best, min(w8[k], d[(k + 8) & 15]));
  return best;
}
__device__ __forceinline__ int fast_measure_ladder(const uint8_t* p, int stride, int min_th) {
  const int v = p[0];
  int ring[16];
  ring[0] = p[3 * stride]; ring[8] = p[-3 * stride]; ring[4] = p[3]; ring[12] = p[-3];
  // every arc of 9 contains pixel 0 or 8, and pixel 4 or 12: most pixels are rejected after 5 loads
  if ((abs(ring[0] - v) <= min_th && abs(ring[8] - v) <= min_th) || (abs(ring[4] - v) <= min_th && abs(ring[12] - v) <= min_th)) return 0;
  ring[1] = p[3 * stride + 1]; ring[2] = p[2 * stride + 2]; ring[3] = p[stride + 3];
  ring[5] = p[-stride + 3]; ring[6] = p[-2 * stride + 2]; ring[7] = p[-3 * stride + 1];
  ring[9] = p[-3 * stride - 1]; ring[10] = p[-2 * stride - 2]; ring[11] = p[-stride - 3];
  ring[13] = p[stride - 3]; ring[14] = p[2 * stride - 2]; ring[15] = p[3 * stride - 1];
  int db[16], dd[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { db[k] = ring[k] - v; dd[k] = v - ring[k]; }
  const int m = max(ring_max_of_window9_min(db), ring_max_of_window9_min(dd));
  return m > min_th ? m : 0;
}

template <bool LADDER>
__global__ void fast_score_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ score, size_t img_bytes, size_t off, int w, int h, int min_th) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const size_t base = (size_t)blockIdx.z * img_bytes + off;
  int m = 0;
  if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
    const uint8_t* p = pyr + base + (size_t)y * w + x;
    m = LADDER ? fast_measure_ladder(p, w, min_th) : fast_measure_bisect(p, w, min_th);
  }
  score[base + (size_t)y * w + x] = (uint8_t)min(max(m, 0), 255);
}

// ---------------------------------------------------------------------------------------------------
// one CTA per cell: cv::FAST(cellROI, iniTh, nonmax) and, if that returns nothing, again with minTh
// (src/ORBextractor.cc:810-817). NMS is cell-local: scores outside the ROI's detection interior are 0.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cell_nms_kernel(const uint8_t* __restrict__ score, size_t img_bytes, const LevelInfo* __restrict__ Ld, int level,
                                                       const CellRect* __restrict__ cells, int total_cells, int ini_th, int min_th,
                                                       uint32_t* __restrict__ slots, int* __restrict__ cell_count, int* __restrict__ err) {
  constexpr int SW = 64, SH = 64;  // interior (wCell x hCell, <= 62 x 62) plus a zero ring
  __shared__ uint8_t sm[SH][SW];
  __shared__ int warp_cnt[4];
  const LevelInfo li = Ld[level];
  const int cell = blockIdx.x, img = blockIdx.y;
  const CellRect cr = cells[li.cell_base + cell];
  const int ix0 = cr.x0 + 3, iy0 = cr.y0 + 3, dw = cr.x1 - 3 - ix0, dh = cr.y1 - 3 - iy0;
  int* out_count = cell_count + (size_t)img * total_cells + li.cell_base + cell;
  if (dw <= 0 || dh <= 0) { if (threadIdx.x == 0) *out_count = 0; return; }
  if (dw > SW - 2 || dh > SH - 2) { if (threadIdx.x == 0) { atomicExch(err, 1); *out_count = 0; } return; }
  const uint8_t* sp = score + (size_t)img * img_bytes + li.plane_off;
  // only the (dh + 2) x (dw + 2) part of the tile that the cell uses (a 30-px cell: ~1000 of the 4096 entries), a warp per row
  for (int yy = threadIdx.x >> 5; yy < dh + 2; yy += 4)
    for (int xx = threadIdx.x & 31; xx < dw + 2; xx += 32) {
      int m = 0;
      if (xx >= 1 && xx <= dw && yy >= 1 && yy <= dh) m = sp[(size_t)(iy0 + yy - 1) * li.w + ix0 + xx - 1];
      sm[yy][xx] = (uint8_t)m;
    }
  uint32_t* my_slots = slots + ((size_t)img * total_cells + li.cell_base + cell) * CELL_CAP;
  const int npix = dw * dh;
  __syncthreads();
  // cv::FAST keeps a corner whose score s = m - 1 (m > th) exceeds the scores of its 8 neighbours, where a neighbour below the
  // threshold scores 0. Since m > th >= any such neighbour's measure, that is: m > th and m greater than all 8 neighbouring measures —
  // the second condition does not depend on the threshold, so no thresholded copy of the tile is needed for either pass.
  for (int pass = 0; pass < 2; ++pass) {
    const int th = pass == 0 ? ini_th : min_th;
    int base = 0;
    for (int c0 = 0; c0 < npix; c0 += 128) {   // ordered (row-major) compaction, 128 pixels at a time
      const int e = c0 + threadIdx.x;
      bool keep = false; int xx = 0, yy = 0, s = 0;
      if (e < npix) {
        yy = e / dw; xx = e - yy * dw;
        const uint8_t* c = &sm[yy + 1][xx + 1];
        const int m = c[0];
        s = m - 1;
        keep = m > th && m > c[-1] && m > c[1] && m > c[-SW - 1] && m > c[-SW] && m > c[-SW + 1] && m > c[SW - 1] && m > c[SW] && m > c[SW + 1];
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
      if (lane == 0) warp_cnt[w] = __popc(bal);
      __syncthreads();
      int off = base;
      for (int k = 0; k < w; ++k) off += warp_cnt[k];
      const int total = warp_cnt[0] + warp_cnt[1] + warp_cnt[2] + warp_cnt[3];
      if (keep) {
        const int idx = off + __popc(bal & ((1u << lane) - 1u));
        if (idx < CELL_CAP) {
          const int xr = ix0 + xx - cr.x0 + cr.offx, yr = iy0 + yy - cr.y0 + cr.offy;   // pt + (j*wCell, i*hCell)
          my_slots[idx] = ((uint32_t)xr << 18) | ((uint32_t)yr << 8) | (uint32_t)s;
        } else atomicExch(err, 2);
      }
      base += total;
      __syncthreads();
    }
    if (base > 0 || pass == 1) {
      if (threadIdx.x == 0) *out_count = min(base, CELL_CAP);
      return;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// quad-tree distribution, one CTA per (level, image) — DistributeOctTree (src/ORBextractor.cc:540-764).
// The std::list of the reference is represented by creation sequence numbers: nodes pushed to the
// front appear in descending creation order, followed by the initial nodes; `order[]` holds the live
// node ids in list order. Ties of the (size, pointer) sort are broken by creation order (see oracle).
// ---------------------------------------------------------------------------------------------------
struct DistSmem {
  short ulx[MAXID], uly[MAXID], urx[MAXID], bly[MAXID];
  int cnt[MAXID];
  int seq[MAXID];
  int childcnt[MAXID][4];
  unsigned short childid[MAXID][4];
  unsigned int best[MAXID];
  unsigned char flags[MAXID];     // bit0 alive, bit1 noMore, bit2 selected for division in this pass
  unsigned short order[MAXID], order2[MAXID], vsize[MAXID], vprev[MAXID], freelist[MAXID];
  int cell_off[MAXCELLS + 1];
  int ctl[8];
  int wsum[3][32];                // per-warp totals of the block scans
};

// Candidates of one (image, level) cached in shared memory behind DistSmem when they fit: packed key point, node id, quadrant. Every
// pass of the quad-tree loop sweeps all candidates two or three times; from global memory each sweep is a chain of dependent loads
// (cell count -> slot -> node id), the kernel's long-scoreboard stall (ncu: 9 % issue active, one CTA per SM).
constexpr int KCACHE = 12288;
constexpr size_t DIST_SMEM_BYTES = ((sizeof(DistSmem) + 15) & ~(size_t)15) + (size_t)KCACHE * 7;

// exclusive scan of up to three counters over the threads of the CTA (<= 32 warps; warp shuffles + one shared-memory hop);
// tot[k] = block totals. Ends with a barrier, so wsum is reusable right away.
template <int NV>
__device__ __forceinline__ void block_exscan(const int (&v)[NV], int (&e)[NV], int (&tot)[NV], int (*wsum)[32]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) inc[k] = v[k];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
#pragma unroll
    for (int k = 0; k < NV; ++k) { const int t = __shfl_up_sync(0xffffffffu, inc[k], o); if (lane >= o) inc[k] += t; }
  if (lane == 31)
#pragma unroll
    for (int k = 0; k < NV; ++k) wsum[k][w] = inc[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int b = 0, t = 0;
#pragma unroll
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) { const int x = wsum[k][j]; if (j < w) b += x; t += x; }
    e[k] = b + inc[k] - v[k]; tot[k] = t;
  }
  __syncthreads();
}

__device__ __forceinline__ int quadrant_of(float x, float y, int mx, int my) {
  if (x < (float)mx) return (y < (float)my) ? 0 : 2;
  return (y < (float)my) ? 1 : 3;
}

constexpr int DIST_THREADS = 1024;   // one CTA per SM (shared memory): 32 warps keep 4x the loads of the candidate sweeps in flight
__global__ void __launch_bounds__(DIST_THREADS) distribute_kernel(const LevelInfo* __restrict__ Ld, int total_cells, const uint32_t* __restrict__ slots,
                                                         const int* __restrict__ cell_count, unsigned short* __restrict__ node_of,
                                                         uint8_t* __restrict__ kq, uint32_t* __restrict__ sel, int* __restrict__ sel_count,
                                                         int sel_cap, int nlevels, int* __restrict__ err, int level0) {
  extern __shared__ unsigned char dist_raw[];
  DistSmem& S = *reinterpret_cast<DistSmem*>(dist_raw);
  const int level = level0 + blockIdx.x, img = blockIdx.y;
  const LevelInfo li = Ld[level];
  const int N = li.nfeat;
  const int tid = threadIdx.x, nthreads = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int* ccount = cell_count + (size_t)img * total_cells + li.cell_base;
  const uint32_t* cslots = slots + ((size_t)img * total_cells + li.cell_base) * CELL_CAP;
  unsigned short* nof = node_of + ((size_t)img * total_cells + li.cell_base) * CELL_CAP;
  uint8_t* kqq = kq + ((size_t)img * total_cells + li.cell_base) * CELL_CAP;
  int* out_count = sel_count + (size_t)img * nlevels + level;
  uint32_t* out = sel + ((size_t)img * nlevels + level) * sel_cap;
  const int ncells = li.ncells;
  // candidate index = cell_off[cell] + slot  (vToDistributeKeys order)
  {   // exclusive scan of the per-cell counts (block scan; a serial walk was ~300 dependent global loads on one thread)
    int base = 0;
    for (int c0 = 0; c0 < ncells; c0 += nthreads) {
      const int c = c0 + tid;
      int v[1] = {c < ncells ? ccount[c] : 0}, e[1], tot[1];
      block_exscan<1>(v, e, tot, S.wsum);
      if (c < ncells) S.cell_off[c] = base + e[0];
      base += tot[0];
    }
    if (tid == 0) S.cell_off[ncells] = base;
  }
  __syncthreads();
  const int M = S.cell_off[ncells];
  if (M == 0) { if (tid == 0) *out_count = 0; return; }
  const int Wd = li.maxBX - li.minBX, Hd = li.maxBY - li.minBY;
  const int nIni = (int)roundf((float)Wd / (float)Hd);
  const float hX = (float)Wd / (float)nIni;
  // ---- initial nodes ----
  for (int i = tid; i < MAXID; i += nthreads) { S.flags[i] = 0; S.cnt[i] = 0; }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < nIni; ++i) {
      S.ulx[i] = (short)(int)(hX * (float)i); S.uly[i] = 0; S.urx[i] = (short)(int)(hX * (float)(i + 1)); S.bly[i] = (short)Hd;
      S.seq[i] = -i; S.flags[i] = 1;
    }
    S.ctl[0] = MAXID - nIni;   // free list size
    S.ctl[1] = 0;              // creation counter
  }
  for (int k = tid; k < MAXID - nIni; k += nthreads) S.freelist[k] = (unsigned short)(MAXID - 1 - k);   // pop from the end -> ids nIni, nIni+1, ...
  __syncthreads();
  unsigned char* cache_base = dist_raw + ((sizeof(DistSmem) + 15) & ~(size_t)15);
  uint32_t* s_pk = reinterpret_cast<uint32_t*>(cache_base);
  unsigned short* s_nof = reinterpret_cast<unsigned short*>(cache_base + (size_t)KCACHE * 4);
  uint8_t* s_kq = cache_base + (size_t)KCACHE * 6;
  const bool cached = M <= KCACHE;
  if (cached) {   // one pass over the cells fills the cache in vToDistributeKeys order (index = cell_off[cell] + slot)
    for (int c = warp; c < ncells; c += nwarps) {
      const int cn = ccount[c], o0 = S.cell_off[c];
      for (int sl = lane; sl < cn; sl += 32) s_pk[o0 + sl] = cslots[c * CELL_CAP + sl];
    }
    __syncthreads();
  }
  // body(packed key point, order index, node id slot, quadrant slot, id to publish as the node's winner)
  auto each_kp = [&](auto body) {
    if (cached) {
      for (int k = tid; k < M; k += nthreads) body(s_pk[k], k, s_nof + k, s_kq + k, k);
    } else {
      for (int c = warp; c < ncells; c += nwarps) {
        const int cn = ccount[c];
        for (int sl = lane; sl < cn; sl += 32) {
          const int kidx = c * CELL_CAP + sl;
          body(cslots[kidx], S.cell_off[c] + sl, nof + kidx, kqq + kidx, kidx);
        }
      }
    }
  };
#define KP_X(pk) ((float)((pk) >> 18))
#define KP_Y(pk) ((float)(((pk) >> 8) & 1023u))
#define KP_RESP(pk) ((int)((pk) & 255u))
  each_kp([&](uint32_t pk, int, unsigned short* pn, uint8_t*, int) {
    const int nid = (int)(KP_X(pk) / hX);
    *pn = (unsigned short)nid;
    atomicAdd(&S.cnt[nid], 1);
  });
  __syncthreads();
  if (tid == 0) {
    int alive = 0;
    for (int i = 0; i < nIni; ++i) {
      if (S.cnt[i] == 0) { S.flags[i] = 0; S.freelist[S.ctl[0]++] = (unsigned short)i; continue; }
      if (S.cnt[i] == 1) S.flags[i] |= 2;
      S.order[alive++] = (unsigned short)i;
    }
    S.ctl[2] = alive;   // live nodes
    S.ctl[3] = 0;       // finish flag
    S.ctl[4] = 0;       // mode: 0 coarse, 1 fine
    S.ctl[5] = 0;       // vsize length
    S.ctl[6] = 0;       // error
  }
  __syncthreads();
  // ---- main loop ----
  for (int guard = 0; guard < 64; ++guard) {
    if (S.ctl[3]) break;
    const int mode = S.ctl[4];
    const int alive = S.ctl[2];
    // 1. select the nodes to divide in this pass
    if (mode == 0) {
      for (int i = tid; i < alive; i += nthreads) { const int id = S.order[i]; if (!(S.flags[id] & 2)) S.flags[id] |= 4; }
    } else {
      // vprev = vsize sorted ascending by (cnt, seq) — the reference's sort of (size, node) pairs —, processed from the end. (cnt, seq) is a
      // total order (seq is unique), so every entry's position is the number of smaller keys: nv^2 / 256 comparisons per thread
      // instead of one thread's insertion sort (nv ~ 100-200 near the feature limit: ~1 M dependent shared-memory cycles, the bulk of
      // this kernel's run time on the fine levels).
      const int nv = S.ctl[5];
      for (int i = tid; i < nv; i += nthreads) {
        const unsigned short v = S.vsize[i];
        const int ci = S.cnt[v], si = S.seq[v];
        int rank = 0;
        for (int j = 0; j < nv; ++j) { const unsigned short u = S.vsize[j]; const int cj = S.cnt[u]; rank += (cj < ci) || (cj == ci && S.seq[u] < si); }
        S.vprev[rank] = v;
      }
      __syncthreads();
      for (int i = tid; i < nv; i += nthreads) S.flags[S.vprev[i]] |= 4;
    }
    __syncthreads();
    for (int i = tid; i < MAXID; i += nthreads)
      if (S.flags[i] & 4) { S.childcnt[i][0] = S.childcnt[i][1] = S.childcnt[i][2] = S.childcnt[i][3] = 0; }
    __syncthreads();
    // 2. child counts of every selected node
    each_kp([&](uint32_t pk, int, unsigned short* pn, uint8_t* pq, int) {
      const int nid = *pn;
      if (S.flags[nid] & 4) {
        const int halfX = (int)ceilf((float)(S.urx[nid] - S.ulx[nid]) / 2), halfY = (int)ceilf((float)(S.bly[nid] - S.uly[nid]) / 2);
        const int q = quadrant_of(KP_X(pk), KP_Y(pk), S.ulx[nid] + halfX, S.uly[nid] + halfY);
        *pq = (uint8_t)q;
        atomicAdd(&S.childcnt[nid][q], 1);
      }
    });
    __syncthreads();
    // 3. list surgery. Coarse mode (every dividable node splits): all threads, three block scans over the list give each
    // node its creation rank (children are numbered in list order, quadrant order), the slot of its big children in vsize
    // and the slot of an untouched node in the kept list — the same lists the sequential walk of the reference builds.
    if (mode == 0) {
      const int nfree0 = S.ctl[0], seqc0 = S.ctl[1], live0 = S.ctl[2];
      int* created = reinterpret_cast<int*>(S.best);   // scratch (best[] is only used at the very end)
      int base_n = 0, base_big = 0, base_kept = 0;
      for (int c0 = 0; c0 < live0; c0 += nthreads) {
        const int i = c0 + tid;
        int id = 0; bool sel = false;
        int v[3] = {0, 0, 0}, e[3], tot[3];
        if (i < live0) {
          id = S.order[i]; sel = (S.flags[id] & 4) != 0;
          if (sel) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int cc = S.childcnt[id][q]; v[0] += cc > 0; v[1] += cc > 1; }
          } else v[2] = 1;
        }
        block_exscan<3>(v, e, tot, S.wsum);
        if (i < live0) {
          if (sel) {
            int r = base_n + e[0], vb = base_big + e[1];
            const int halfX = (int)ceilf((float)(S.urx[id] - S.ulx[id]) / 2), halfY = (int)ceilf((float)(S.bly[id] - S.uly[id]) / 2);
            const int mx = S.ulx[id] + halfX, my = S.uly[id] + halfY;
            for (int q = 0; q < 4; ++q) {
              const int cc = S.childcnt[id][q];
              S.childid[id][q] = 0xFFFF;
              if (cc == 0) continue;
              if (r >= nfree0) { S.ctl[6] = 3; ++r; continue; }
              const int ch = S.freelist[nfree0 - 1 - r];
              S.ulx[ch] = (short)((q & 1) ? mx : S.ulx[id]); S.urx[ch] = (short)((q & 1) ? S.urx[id] : mx);
              S.uly[ch] = (short)((q & 2) ? my : S.uly[id]); S.bly[ch] = (short)((q & 2) ? S.bly[id] : my);
              S.cnt[ch] = cc; S.seq[ch] = seqc0 + r + 1; S.flags[ch] = (unsigned char)(1 | (cc == 1 ? 2 : 0));
              S.childid[id][q] = (unsigned short)ch;
              created[r] = ch;
              if (cc > 1) S.vsize[vb++] = (unsigned short)ch;
              ++r;
            }
          } else S.order2[base_kept + e[2]] = (unsigned short)id;
        }
        base_n += tot[0]; base_big += tot[1]; base_kept += tot[2];
      }
      __syncthreads();
      const int ncreated = min(base_n, nfree0);
      // new list = children in reverse creation order, then the untouched (noMore) nodes in their old order
      for (int p = tid; p < ncreated; p += nthreads) S.order[p] = (unsigned short)created[ncreated - 1 - p];
      for (int p = tid; p < base_kept; p += nthreads) S.order[ncreated + p] = S.order2[p];
      __syncthreads();
      if (tid == 0) {
        const int live = ncreated + base_kept;
        S.ctl[0] = nfree0 - ncreated; S.ctl[1] = seqc0 + ncreated; S.ctl[2] = live; S.ctl[5] = base_big;
        if (live >= N || live == live0) S.ctl[3] = 1;
        else if (live + base_big * 3 > N) S.ctl[4] = 1;
      }
    } else if (tid == 0) {   // fine mode (near the feature limit): sequential, the largest nodes first until the limit is reached
      int nfree = S.ctl[0], seqc = S.ctl[1], live = S.ctl[2];
      auto make_children = [&](int id, int* created, int& ncreated, int* nToExpand, int& nvs) {
        const int halfX = (int)ceilf((float)(S.urx[id] - S.ulx[id]) / 2), halfY = (int)ceilf((float)(S.bly[id] - S.uly[id]) / 2);
        const int mx = S.ulx[id] + halfX, my = S.uly[id] + halfY;
        for (int q = 0; q < 4; ++q) {
          const int cc = S.childcnt[id][q];
          S.childid[id][q] = 0xFFFF;
          if (cc == 0) continue;
          if (nfree == 0) { S.ctl[6] = 3; continue; }
          const int ch = S.freelist[--nfree];
          S.ulx[ch] = (short)((q & 1) ? mx : S.ulx[id]); S.urx[ch] = (short)((q & 1) ? S.urx[id] : mx);
          S.uly[ch] = (short)((q & 2) ? my : S.uly[id]); S.bly[ch] = (short)((q & 2) ? S.bly[id] : my);
          S.cnt[ch] = cc; S.seq[ch] = ++seqc; S.flags[ch] = (unsigned char)(1 | (cc == 1 ? 2 : 0));
          S.childid[id][q] = (unsigned short)ch;
          created[ncreated++] = ch;
          if (cc > 1) { if (nToExpand) ++*nToExpand; S.vsize[nvs++] = (unsigned short)ch; }
        }
      };
      int* created = reinterpret_cast<int*>(S.best);   // scratch (best[] is only used at the very end)
      int ncreated = 0, nvs = 0;
      {
        const int prevSize = live;
        const int nv = S.ctl[5];
        int processed_from = nv;   // vprev[processed_from .. nv) were divided
        for (int j = nv - 1; j >= 0; --j) {
          const int id = S.vprev[j];
          const int before = ncreated;
          make_children(id, created, ncreated, nullptr, nvs);
          live += (ncreated - before) - 1;
          S.flags[id] |= 8;   // divided in this pass
          processed_from = j;
          if (live >= N) break;
        }
        // selected but not reached (after the break): keep them alive and un-divided
        for (int j = 0; j < processed_from; ++j) S.flags[S.vprev[j]] &= (unsigned char)~4;
        // list: children (reverse creation) in front, then the old list without the divided nodes
        int nkept = 0;
        const int old_live = S.ctl[2];
        for (int i = 0; i < old_live; ++i) { const int id = S.order[i]; if (!(S.flags[id] & 8)) S.order2[nkept++] = (unsigned short)id; }
        int p = 0;
        for (int i = ncreated - 1; i >= 0; --i) S.order[p++] = (unsigned short)created[i];
        for (int i = 0; i < nkept; ++i) S.order[p++] = S.order2[i];
        live = p;
        S.ctl[5] = nvs;
        if (live >= N || live == prevSize) S.ctl[3] = 1;
      }
      S.ctl[0] = nfree; S.ctl[1] = seqc; S.ctl[2] = live;
    }
    __syncthreads();
    // 4. move the candidates of divided nodes to their children, then recycle the parents
    each_kp([&](uint32_t, int, unsigned short* pn, uint8_t* pq, int) {
      const int nid = *pn;
      const int fl = S.flags[nid];
      if ((fl & 4) && (mode == 0 || (fl & 8))) *pn = S.childid[nid][*pq];
    });
    __syncthreads();
    {   // recycle the divided parents: ordered compaction of their ids onto the free list (ascending id, like a serial walk)
      int nfree = S.ctl[0];
      __syncthreads();
      for (int c0 = 0; c0 < MAXID; c0 += nthreads) {
        const int i = c0 + tid;
        const int fl = S.flags[i];
        const bool fr = (fl & 4) && (mode == 0 || (fl & 8));
        if (fr) S.flags[i] = 0;
        else if (fl & 4) S.flags[i] = (unsigned char)(fl & ~4);
        int v[1] = {fr ? 1 : 0}, e[1], tot[1];
        block_exscan<1>(v, e, tot, S.wsum);
        if (fr) S.freelist[nfree + e[0]] = (unsigned short)i;
        nfree += tot[0];
      }
      if (tid == 0) S.ctl[0] = nfree;
      __syncthreads();
    }
  }
  if (S.ctl[6] && tid == 0) atomicExch(err, 3);
  // ---- best candidate per live node: max response, first in insertion order on ties ----
  for (int i = tid; i < MAXID; i += nthreads) S.best[i] = 0u;
  __syncthreads();
  each_kp([&](uint32_t pk, int kord, unsigned short* pn, uint8_t*, int) {
    const unsigned key = ((unsigned)KP_RESP(pk) << 20) | (0xFFFFFu - (unsigned)kord);
    atomicMax(&S.best[*pn], key);
  });
  __syncthreads();
  each_kp([&](uint32_t pk, int kord, unsigned short* pn, uint8_t*, int kid) {
    const int nid = *pn;
    const unsigned key = ((unsigned)KP_RESP(pk) << 20) | (0xFFFFFu - (unsigned)kord);
    if (S.best[nid] == key) S.seq[nid] = kid;   // the winner (unique key) publishes its index; seq[] is free once the tree is final
  });
  __syncthreads();
  const int live = S.ctl[2];
  if (live > sel_cap) { if (tid == 0) { atomicExch(err, 4); *out_count = 0; } return; }
  for (int i = tid; i < live; i += nthreads) { const int bk = S.seq[S.order[i]]; out[i] = cached ? s_pk[bk] : cslots[bk]; }
  if (tid == 0) *out_count = live;
#undef KP_X
#undef KP_Y
#undef KP_RESP
}

// ---------------------------------------------------------------------------------------------------
// cv::GaussianBlur(7x7, sigma=2, BORDER_REFLECT_101) for 8UC1: integer taps, (V + 2^15) >> 16
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; return p; }
constexpr int BLUR_ROWS = 32;   // output rows per CTA: 38 filtered rows for 32 outputs (1.19x) instead of 14 for 8 (1.75x)
__global__ void __launch_bounds__(256) blur7_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, size_t img_bytes, size_t off, int w, int h,
                                                    int t0, int t1, int t2, int t3) {
  __shared__ int hs[BLUR_ROWS + 6][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + tx, y0 = blockIdx.y * BLUR_ROWS;
  const uint8_t* src = pyr + (size_t)blockIdx.z * img_bytes + off;
  const bool interior_x = x >= 3 && x + 3 < w;   // no reflection needed along the row (all but the first / last column tile)
  for (int r = ty; r < BLUR_ROWS + 6; r += 8) {
    const int yr = y0 + r - 3;
    if (yr >= h + 3) break;                       // rows below the image are not needed by any output of this tile
    const int yy = reflect101(yr, h);
    int s = 0;
    if (x < w) {
      const uint8_t* p = src + (size_t)yy * w;
      if (interior_x) s = t0 * (p[x - 3] + p[x + 3]) + t1 * (p[x - 2] + p[x + 2]) + t2 * (p[x - 1] + p[x + 1]) + t3 * p[x];
      else s = t0 * (p[reflect101(x - 3, w)] + p[reflect101(x + 3, w)]) + t1 * (p[reflect101(x - 2, w)] + p[reflect101(x + 2, w)]) +
               t2 * (p[reflect101(x - 1, w)] + p[reflect101(x + 1, w)]) + t3 * p[x];
    }
    hs[r][tx] = s;
  }
  __syncthreads();
  if (x >= w) return;
#pragma unroll
  for (int k = 0; k < BLUR_ROWS / 8; ++k) {
    const int ry = ty + 8 * k, y = y0 + ry;
    if (y < h) {
      const int v = t0 * (hs[ry][tx] + hs[ry + 6][tx]) + t1 * (hs[ry + 1][tx] + hs[ry + 5][tx]) + t2 * (hs[ry + 2][tx] + hs[ry + 4][tx]) + t3 * hs[ry + 3][tx];
      blur[(size_t)blockIdx.z * img_bytes + off + (size_t)y * w + x] = (uint8_t)min(255, max(0, (v + (1 << 15)) >> 16));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// one warp per selected keypoint: IC_Angle (integer moments + cv::fastAtan2), steered BRIEF, KeyPoint
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float k = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k, p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON)); c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON)); c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

__global__ void __launch_bounds__(128) orient_describe_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, size_t img_bytes,
                                                              const LevelInfo* __restrict__ Ld, int nlevels, const uint32_t* __restrict__ sel,
                                                              const int* __restrict__ sel_count, int sel_cap, int out_cap,
                                                              tslam_keypoint* __restrict__ kp_out, uint8_t* __restrict__ desc_out,
                                                              int* __restrict__ counts, int* __restrict__ err) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.y;
  const int level = warp / sel_cap, i = warp - level * sel_cap;
  if (level >= nlevels) return;
  const int* sc = sel_count + (size_t)img * nlevels;
  if (i >= sc[level]) return;
  int offset = 0;
  for (int l = 0; l < level; ++l) offset += sc[l];
  const int o = offset + i;
  if (o >= out_cap) { if (lane == 0) atomicExch(err, 5); return; }
  const LevelInfo li = Ld[level];
  const uint32_t pk = sel[((size_t)img * nlevels + level) * sel_cap + i];
  const int x = (int)(pk >> 18) + li.minBX, y = (int)((pk >> 8) & 1023u) + li.minBY, resp = (int)(pk & 255u);
  const uint8_t* im = pyr + (size_t)img * img_bytes + li.plane_off;
  const uint8_t* bl = blur + (size_t)img * img_bytes + li.plane_off;
  // IC_Angle (src/ORBextractor.cc:77-108), integer moments: lane <-> COLUMN u = lane - 15, loop over the rows, so that one load
  // instruction reads 31 neighbouring bytes of one image row (1-2 sectors; with lane <-> row it touched 31 rows = 31 sectors and the
  // kernel was bound by LSU wavefronts). m10 = sum_u u * (column sum), m01 = sum_u (sum_v v * pixel): exact in integers.
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int u = lane - 15, au = abs(u);
    const uint8_t* col = im + (size_t)y * li.w + x + u;
    int cs = 0, vs = 0;
#pragma unroll
    for (int v = -15; v <= 15; ++v) {
      const int val = au <= c_umax[v < 0 ? -v : v] ? (int)col[v * li.w] : 0;   // uniform index: a constant-cache broadcast
      cs += val; vs += v * val;
    }
    m10 = u * cs; m01 = vs;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, s); m01 += __shfl_xor_sync(0xffffffffu, m01, s); }
  const float angle = fast_atan2_deg((float)m01, (float)m10);
  // steered BRIEF: lane <-> descriptor byte
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  const float ang = __fmul_rn(angle, factorPI);
  double sd, cd;
  tsl_det_sincos((double)ang, &sd, &cd);
  const float a = (float)cd, b = (float)sd;
  const uint8_t* center = bl + (size_t)y * li.w + x;
  int val = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t pw = __ldg(g_pat_packed + 32 * k + lane);
    const float x0 = (float)(int8_t)(pw & 255u), y0 = (float)(int8_t)((pw >> 8) & 255u), x1 = (float)(int8_t)((pw >> 16) & 255u), y1 = (float)(int8_t)(pw >> 24);
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a))), c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a))), c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int t0 = center[r0 * li.w + c0], t1 = center[r1 * li.w + c1];
    val |= (t0 < t1) << k;
  }
  desc_out[((size_t)img * out_cap + o) * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    tslam_keypoint kp;
    float fx = (float)x, fy = (float)y;
    if (level != 0) { fx = __fmul_rn(fx, li.scale); fy = __fmul_rn(fy, li.scale); }
    kp.x = fx; kp.y = fy; kp.size = (float)li.patch; kp.angle = angle; kp.response = (float)resp; kp.octave = level; kp.class_id = -1;
    kp_out[(size_t)img * out_cap + o] = kp;
  }
}

__global__ void counts_kernel(const int* __restrict__ sel_count, int nlevels, int out_cap, int* __restrict__ counts, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int t = 0;
  for (int l = 0; l < nlevels; ++l) t += sel_count[(size_t)i * nlevels + l];
  counts[i] = min(t, out_cap);
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
static void resize_table(int sn, int dn, std::vector<int>& ofs, std::vector<short>& ab) {
  const double scale = (double)sn / dn;
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    ofs.push_back(s);
    ab.push_back((short)cv_round_host((1.f - f) * 2048.f));
    ab.push_back((short)cv_round_host(f * 2048.f));
  }
}

static int orb_configure(tslam_orb* o, int w, int h, int n_imgs) {
  tslam_ctx* ctx = o->ctx; cudaStream_t st = ctx->stream;
  if (o->w != w || o->h != h) {
    o->w = 0; o->h = 0; o->n_alloc = 0;   // committed only after every check below has passed: a failed call must not leave a half-built table behind a matching size
    o->L.assign(o->nlevels, LevelInfo());
    std::vector<CellRect> cells;
    std::vector<int> xofs, yofs; std::vector<short> xa, ya;
    o->xtab_off.assign(o->nlevels, 0); o->ytab_off.assign(o->nlevels, 0);
    size_t off = 0; int cell_base = 0;
    for (int l = 0; l < o->nlevels; ++l) {
      LevelInfo& li = o->L[l];
      const float s = o->mvInvScale[l];
      li.w = cv_round_host((float)w * s); li.h = cv_round_host((float)h * s);
      if (li.w < 2 * ORB_EDGE + 8 || li.h < 2 * ORB_EDGE + 8) return set_error(TSLAM_ERR_ARG, "image %dx%d too small for %d levels", w, h, o->nlevels);
      li.minBX = ORB_EDGE - 3; li.minBY = li.minBX; li.maxBX = li.w - ORB_EDGE + 3; li.maxBY = li.h - ORB_EDGE + 3;
      if (li.maxBX - li.minBX > 1023 || li.maxBY - li.minBY > 1023) return set_error(TSLAM_ERR_ARG, "image %dx%d exceeds the 1039-px packing limit", w, h);
      li.nfeat = o->perLevel[l]; li.scale = o->mvScale[l]; li.patch = (int)(ORB_PATCH * o->mvScale[l]);
      li.plane_off = off; off += (size_t)li.w * li.h; off = (off + 15) & ~(size_t)15;
      // cells exactly as ComputeKeyPointsOctTree (src/ORBextractor.cc:779-807)
      const float W = 30;
      const float width = (float)(li.maxBX - li.minBX), height = (float)(li.maxBY - li.minBY);
      const int nCols = (int)(width / W), nRows = (int)(height / W);
      if (nCols < 1 || nRows < 1) return set_error(TSLAM_ERR_ARG, "level %d too small", l);
      if ((int)std::round(width / height) < 1) return set_error(TSLAM_ERR_ARG, "aspect ratio w/h < 0.5 is not supported (the reference divides by zero there)");
      const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
      li.cell_base = cell_base; li.ncells = 0;
      for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(li.minBY + i * hCell); float maxY = iniY + hCell + 6;
        if (iniY >= li.maxBY - 3) continue;
        if (maxY > li.maxBY) maxY = (float)li.maxBY;
        for (int j = 0; j < nCols; ++j) {
          const float iniX = (float)(li.minBX + j * wCell); float maxX = iniX + wCell + 6;
          if (iniX >= li.maxBX - 6) continue;
          if (maxX > li.maxBX) maxX = (float)li.maxBX;
          cells.push_back(CellRect{(short)(int)iniX, (short)(int)iniY, (short)(int)maxX, (short)(int)maxY, (short)(j * wCell), (short)(i * hCell)});
          ++li.ncells;
        }
      }
      if (li.ncells > MAXCELLS) return set_error(TSLAM_ERR_ARG, "too many cells (%d) at level %d", li.ncells, l);
      if (li.nfeat + 8 > MAXID / 2) return set_error(TSLAM_ERR_ARG, "nfeatures per level %d exceeds the node budget", li.nfeat);
      cell_base += li.ncells;
      if (l > 0) {
        o->xtab_off[l] = (int)xofs.size(); o->ytab_off[l] = (int)yofs.size();
        resize_table(o->L[l - 1].w, li.w, xofs, xa);
        resize_table(o->L[l - 1].h, li.h, yofs, ya);
      }
    }
    o->total_cells = cell_base; o->img_bytes = off;
    int mx = 0, tot = 0;
    for (int l = 0; l < o->nlevels; ++l) { mx = std::max(mx, o->perLevel[l] + 4); tot += o->perLevel[l] + 4; }
    o->sel_cap = mx; o->out_cap = tot;
    TSL_CUDA(o->cells.upload(cells.data(), cells.size(), st));
    TSL_CUDA(o->Ld.upload(o->L.data(), o->L.size(), st));
    TSL_CUDA(o->xofs.upload(xofs.data(), xofs.size(), st)); TSL_CUDA(o->yofs.upload(yofs.data(), yofs.size(), st));
    TSL_CUDA(o->xa.upload(xa.data(), xa.size(), st)); TSL_CUDA(o->ya.upload(ya.data(), ya.size(), st));
    TSL_CUDA(cudaStreamSynchronize(st));
    o->w = w; o->h = h;
  }
  if (n_imgs > o->n_alloc) {
    const size_t n = n_imgs;
    TSL_CUDA(o->pyr.reserve(n * o->img_bytes)); TSL_CUDA(o->blur.reserve(n * o->img_bytes)); TSL_CUDA(o->score.reserve(n * o->img_bytes));
    TSL_CUDA(o->slots.reserve(n * o->total_cells * CELL_CAP)); TSL_CUDA(o->node_of.reserve(n * o->total_cells * CELL_CAP));
    TSL_CUDA(o->kq.reserve(n * o->total_cells * CELL_CAP));
    TSL_CUDA(o->cell_count.reserve(n * o->total_cells)); TSL_CUDA(o->sel_count.reserve(n * o->nlevels));
    TSL_CUDA(o->sel.reserve(n * o->nlevels * o->sel_cap));
    TSL_CUDA(o->kp.reserve(n * o->out_cap)); TSL_CUDA(o->desc.reserve(n * o->out_cap * 32)); TSL_CUDA(o->counts.reserve(n));
    TSL_CUDA(o->err.reserve(1));
    o->n_alloc = n_imgs;
  }
  return TSLAM_OK;
}

// the whole extractor on images already stored as level 0 of the pyramid records
static int orb_run(tslam_orb* o, int n) {
  tslam_ctx* ctx = o->ctx; cudaStream_t st = ctx->stream;
  // the opt-in is per device: tracked per context, not per process (several contexts on several GPUs may live in one process)
  if (!ctx->attr_orb) { TSL_CUDA(cudaFuncSetAttribute(distribute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIST_SMEM_BYTES)); ctx->attr_orb = true; }
  TSL_CUDA(cudaMemsetAsync(o->err.p, 0, sizeof(int), st));
  static const int T0[4] = {18, 34, 48, 56}, T1[4] = {18, 34, 49, 55};
  const int* T = o->blur_variant == 1 ? T1 : T0;
  // Two streams. The quad-tree distribution of a level (64 CTAs of 8 warps holding 115 KB of shared memory each: latency bound, one
  // per SM) starts on the second stream as soon as that level's candidates exist, and runs beside the FAST / NMS kernels of the next
  // levels and beside the 7x7 Gaussian (which only the descriptors need); the streams join before orient_describe_kernel.
  static const bool overlap = [] { const char* e = getenv("TSLAM_ORB_OVERLAP"); return !(e && e[0] == '0'); }();
  if (overlap && !o->s2) {
    TSL_CUDA(cudaStreamCreateWithFlags(&o->s2, cudaStreamNonBlocking));
    TSL_CUDA(cudaEventCreateWithFlags(&o->ev_fork, cudaEventDisableTiming)); TSL_CUDA(cudaEventCreateWithFlags(&o->ev_join, cudaEventDisableTiming));
    for (int l = 0; l < 16; ++l) { TSL_CUDA(cudaEventCreateWithFlags(&o->ev_level[l], cudaEventDisableTiming)); TSL_CUDA(cudaEventCreateWithFlags(&o->ev_res[l], cudaEventDisableTiming)); }
  }
  if (o->nlevels > 16) return set_error(TSLAM_ERR_ARG, "more than 16 pyramid levels");
  cudaStream_t sd = overlap ? o->s2 : st;
  // the pyramid (each level from the previous one: seven small dependent launches) on the second stream, while the first one already
  // scores level 0; a level's FAST pass waits for that level's resize only
  if (overlap) { TSL_CUDA(cudaEventRecord(o->ev_fork, st)); TSL_CUDA(cudaStreamWaitEvent(sd, o->ev_fork, 0)); }
  for (int l = 1; l < o->nlevels; ++l) {
    const LevelInfo& s = o->L[l - 1]; const LevelInfo& d = o->L[l];
    dim3 grid((d.w + 127) / 128, d.h, n);
    LAUNCH(resize_kernel<<<grid, 128, 0, sd>>>(o->pyr.p, o->pyr.p, o->img_bytes, s.plane_off, d.plane_off, s.w, s.h, d.w, d.h,
                                                o->xofs.p + o->xtab_off[l], o->xa.p + 2 * o->xtab_off[l], o->yofs.p + o->ytab_off[l], o->ya.p + 2 * o->ytab_off[l]));
    if (overlap) TSL_CUDA(cudaEventRecord(o->ev_res[l], sd));
  }
  for (int l = 0; l < o->nlevels; ++l) {
    const LevelInfo& li = o->L[l];
    dim3 grid((li.w + 127) / 128, li.h, n);
    if (overlap && l > 0) TSL_CUDA(cudaStreamWaitEvent(st, o->ev_res[l], 0));
    static const bool bisect = getenv("TSLAM_FAST_BISECT") != nullptr;   // the first (bisection) formulation, kept for A/B runs
    if (bisect) LAUNCH(fast_score_kernel<false><<<grid, 128, 0, st>>>(o->pyr.p, o->score.p, o->img_bytes, li.plane_off, li.w, li.h, o->minTh));
    else LAUNCH(fast_score_kernel<true><<<grid, 128, 0, st>>>(o->pyr.p, o->score.p, o->img_bytes, li.plane_off, li.w, li.h, o->minTh));
    LAUNCH(cell_nms_kernel<<<dim3(li.ncells, n), 128, 0, st>>>(o->score.p, o->img_bytes, o->Ld.p, l, o->cells.p, o->total_cells, o->iniTh, o->minTh,
                                                                 o->slots.p, o->cell_count.p, o->err.p));
    // two groups: the two finest levels (most candidates, the longest quad-tree loops) as soon as they are ready — they run beside
    // the FAST / NMS passes of the coarser levels —, the rest beside the Gaussian. (One launch per level serialises eight
    // latency-bound 64-CTA grids on the second stream: 3.46 ms per batch; one launch for all levels after the last NMS: 2.84 ms.)
    const int split = o->nlevels > 2 ? 2 : o->nlevels;
    if (l == split - 1 || l == o->nlevels - 1) {
      const int l0 = l == split - 1 ? 0 : split;
      if (overlap) { TSL_CUDA(cudaEventRecord(o->ev_level[l], st)); TSL_CUDA(cudaStreamWaitEvent(sd, o->ev_level[l], 0)); }
      if (l - l0 + 1 > 0)
        LAUNCH(distribute_kernel<<<dim3(l - l0 + 1, n), DIST_THREADS, DIST_SMEM_BYTES, sd>>>(o->Ld.p, o->total_cells, o->slots.p, o->cell_count.p, o->node_of.p, o->kq.p,
                                                                                      o->sel.p, o->sel_count.p, o->sel_cap, o->nlevels, o->err.p, l0));
    }
  }
  for (int l = 0; l < o->nlevels; ++l) {
    const LevelInfo& li = o->L[l];
    dim3 bgrid((li.w + 31) / 32, (li.h + BLUR_ROWS - 1) / BLUR_ROWS, n);
    LAUNCH(blur7_kernel<<<bgrid, 256, 0, st>>>(o->pyr.p, o->blur.p, o->img_bytes, li.plane_off, li.w, li.h, T[0], T[1], T[2], T[3]));
  }
  if (overlap) { TSL_CUDA(cudaEventRecord(o->ev_join, sd)); TSL_CUDA(cudaStreamWaitEvent(st, o->ev_join, 0)); }
  LAUNCH(counts_kernel<<<(n + 127) / 128, 128, 0, st>>>(o->sel_count.p, o->nlevels, o->out_cap, o->counts.p, n));
  TSL_CHECK_LAUNCH();
  const int warps = o->nlevels * o->sel_cap;
  LAUNCH(orient_describe_kernel<<<dim3((warps * 32 + 127) / 128, n), 128, 0, st>>>(o->pyr.p, o->blur.p, o->img_bytes, o->Ld.p, o->nlevels, o->sel.p,
                                                                                    o->sel_count.p, o->sel_cap, o->out_cap, o->kp.p, o->desc.p, o->counts.p, o->err.p));
  TSL_CHECK_LAUNCH();
  o->last_n = n;
  return TSLAM_OK;
}

static int orb_upload(tslam_orb* o, const uint8_t* const* imgs, int n, int w, int h, int stride) {
  cudaStream_t st = o->ctx->stream;
  for (int i = 0; i < n; ++i) {
    if (!imgs[i]) return set_error(TSLAM_ERR_ARG, "image %d is NULL", i);
    TSL_CUDA(cudaMemcpy2DAsync(o->pyr.p + (size_t)i * o->img_bytes, w, imgs[i], stride, w, h, cudaMemcpyHostToDevice, st));
  }
  return TSLAM_OK;
}

static int orb_check_err(tslam_orb* o) {
  int e = 0;
  TSL_CUDA(cudaMemcpyAsync(&e, o->err.p, sizeof(int), cudaMemcpyDeviceToHost, o->ctx->stream));
  TSL_CUDA(cudaStreamSynchronize(o->ctx->stream));
  if (e) return set_error(TSLAM_ERR_NUMERIC, "ORB extractor capacity exceeded (code %d: 1/2 cell, 3 node ids, 4 selection, 5 output)", e);
  return TSLAM_OK;
}

}  // namespace tsl

using namespace tsl;

extern "C" {

int tslam_orb_create(tslam_ctx* ctx, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int blur_variant, tslam_orb** out) {
  if (!ctx || !out) return set_error(TSLAM_ERR_ARG, "null argument");
  if (nlevels < 1 || nlevels > MAX_LEVELS || nfeatures < 1 || !(scale_factor > 1.0f) || min_th < 1 || ini_th < min_th)
    return set_error(TSLAM_ERR_ARG, "bad ORB parameters");
  TSL_CUDA(cudaSetDevice(ctx->device));
  tslam_orb* o = new tslam_orb();
  o->ctx = ctx; o->nfeatures = nfeatures; o->nlevels = nlevels; o->iniTh = ini_th; o->minTh = min_th; o->blur_variant = blur_variant; o->scaleFactor = scale_factor;
  // src/ORBextractor.cc:416-447 (float arithmetic kept as written)
  o->mvScale.resize(nlevels); o->mvInvScale.resize(nlevels);
  o->mvScale[0] = 1.0f;
  for (int i = 1; i < nlevels; ++i) o->mvScale[i] = o->mvScale[i - 1] * scale_factor;
  for (int i = 0; i < nlevels; ++i) o->mvInvScale[i] = 1.0f / o->mvScale[i];
  o->perLevel.resize(nlevels);
  const float factor = 1.0f / scale_factor;
  float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; ++l) { o->perLevel[l] = cv_round_host(nDesired); sum += o->perLevel[l]; nDesired *= factor; }
  o->perLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
  // umax (src/ORBextractor.cc:453-470)
  int v, v0, vmax = (int)std::floor(ORB_HALF_PATCH * std::sqrt(2.f) / 2 + 1);
  const int vmin = (int)std::ceil(ORB_HALF_PATCH * std::sqrt(2.f) / 2);
  const double hp2 = ORB_HALF_PATCH * ORB_HALF_PATCH;
  for (v = 0; v <= vmax; ++v) o->umax[v] = cv_round_host(std::sqrt(hp2 - v * v));
  for (v = ORB_HALF_PATCH, v0 = 0; v >= vmin; --v) { while (o->umax[v0] == o->umax[v0 + 1]) ++v0; o->umax[v] = v0; ++v0; }
  {
    uint32_t packed[8 * 32];
    for (int k = 0; k < 8; ++k)
      for (int ln = 0; ln < 32; ++ln) {
        const int8_t* q = h_pattern + ln * 32 + 4 * k;
        packed[32 * k + ln] = (uint32_t)(uint8_t)q[0] | ((uint32_t)(uint8_t)q[1] << 8) | ((uint32_t)(uint8_t)q[2] << 16) | ((uint32_t)(uint8_t)q[3] << 24);
      }
    TSL_CUDA(cudaMemcpyToSymbol(g_pat_packed, packed, sizeof(packed)));
  }
  TSL_CUDA(cudaMemcpyToSymbol(c_umax, o->umax, sizeof(o->umax)));
  *out = o;
  return TSLAM_OK;
}

void tslam_orb_destroy(tslam_orb* o) {
  if (!o) return;
  cudaSetDevice(o->ctx->device);
  delete o;
}

int tslam_orb_extract(tslam_orb* o, const uint8_t* const* imgs, int n_imgs, int w, int hgt, int stride, int max_kp, tslam_keypoint* kp_out,
                      uint8_t* desc_out, int32_t* counts_out) {
  if (!o || !imgs || !kp_out || !desc_out || !counts_out) return set_error(TSLAM_ERR_ARG, "null argument");
  if (n_imgs <= 0 || stride < w) return set_error(TSLAM_ERR_ARG, "bad image batch");
  TSL_CUDA(cudaSetDevice(o->ctx->device));
  int rc = orb_configure(o, w, hgt, n_imgs);
  if (rc) return rc;
  if (max_kp < o->out_cap) return set_error(TSLAM_ERR_ARG, "max_kp %d < required capacity %d", max_kp, o->out_cap);
  if ((rc = orb_upload(o, imgs, n_imgs, w, hgt, stride))) return rc;
  if ((rc = orb_run(o, n_imgs))) return rc;
  cudaStream_t st = o->ctx->stream;
  TSL_CUDA(cudaMemcpy2DAsync(kp_out, (size_t)max_kp * sizeof(tslam_keypoint), o->kp.p, (size_t)o->out_cap * sizeof(tslam_keypoint),
                             (size_t)o->out_cap * sizeof(tslam_keypoint), n_imgs, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpy2DAsync(desc_out, (size_t)max_kp * 32, o->desc.p, (size_t)o->out_cap * 32, (size_t)o->out_cap * 32, n_imgs, cudaMemcpyDeviceToHost, st));
  TSL_CUDA(cudaMemcpyAsync(counts_out, o->counts.p, sizeof(int) * n_imgs, cudaMemcpyDeviceToHost, st));
  return orb_check_err(o);
}

int tslam_orb_level_size(tslam_orb* o, int level, int* w, int* hgt) {
  if (!o || level < 0 || level >= o->nlevels || o->L.empty()) return set_error(TSLAM_ERR_ARG, "no pyramid yet / bad level");
  *w = o->L[level].w; *hgt = o->L[level].h;
  return TSLAM_OK;
}

int tslam_orb_get_level(tslam_orb* o, int img, int level, uint8_t* out) {
  if (!o || !out || level < 0 || level >= o->nlevels || img < 0 || img >= o->last_n) return set_error(TSLAM_ERR_ARG, "bad image / level");
  TSL_CUDA(cudaSetDevice(o->ctx->device));
  const LevelInfo& li = o->L[level];
  TSL_CUDA(cudaMemcpyAsync(out, o->pyr.p + (size_t)img * o->img_bytes + li.plane_off, (size_t)li.w * li.h, cudaMemcpyDeviceToHost, o->ctx->stream));
  TSL_CUDA(cudaStreamSynchronize(o->ctx->stream));
  return TSLAM_OK;
}

// Intermediate stages of the last extract call, for stage-by-stage parity tests:
//   what = 0: FAST measure plane of `level` (w_l*h_l u8)      1: candidate count of `level` (1 int), then the candidates in
//   vToDistributeKeys order as (x_rel, y_rel, response) int triplets      2: quad-tree winners of `level` (count, then triplets)
int tslam_orb_debug_get(tslam_orb* o, int what, int img, int level, void* out, int out_bytes) {
  if (!o || !out || level < 0 || level >= o->nlevels || img < 0 || img >= o->last_n) return set_error(TSLAM_ERR_ARG, "bad image / level");
  TSL_CUDA(cudaSetDevice(o->ctx->device));
  cudaStream_t st = o->ctx->stream;
  const LevelInfo& li = o->L[level];
  if (what == 0) {
    if (out_bytes < li.w * li.h) return set_error(TSLAM_ERR_ARG, "buffer too small");
    TSL_CUDA(cudaMemcpyAsync(out, o->score.p + (size_t)img * o->img_bytes + li.plane_off, (size_t)li.w * li.h, cudaMemcpyDeviceToHost, st));
    TSL_CUDA(cudaStreamSynchronize(st));
    return TSLAM_OK;
  }
  int* io = (int*)out;
  if (what == 1) {
    std::vector<int> cc(li.ncells);
    std::vector<uint32_t> sl((size_t)li.ncells * CELL_CAP);
    TSL_CUDA(cudaMemcpyAsync(cc.data(), o->cell_count.p + (size_t)img * o->total_cells + li.cell_base, sizeof(int) * li.ncells, cudaMemcpyDeviceToHost, st));
    TSL_CUDA(cudaMemcpyAsync(sl.data(), o->slots.p + ((size_t)img * o->total_cells + li.cell_base) * CELL_CAP, sizeof(uint32_t) * sl.size(), cudaMemcpyDeviceToHost, st));
    TSL_CUDA(cudaStreamSynchronize(st));
    int n = 0;
    for (int c = 0; c < li.ncells; ++c) n += cc[c];
    if (out_bytes < (int)sizeof(int) * (1 + 3 * n)) return set_error(TSLAM_ERR_ARG, "buffer too small (%d candidates)", n);
    io[0] = n; int k = 1;
    for (int c = 0; c < li.ncells; ++c)
      for (int q = 0; q < cc[c]; ++q) { const uint32_t pk = sl[(size_t)c * CELL_CAP + q]; io[k++] = (int)(pk >> 18); io[k++] = (int)((pk >> 8) & 1023u); io[k++] = (int)(pk & 255u); }
    return TSLAM_OK;
  }
  if (what == 2) {
    int n = 0;
    TSL_CUDA(cudaMemcpyAsync(&n, o->sel_count.p + (size_t)img * o->nlevels + level, sizeof(int), cudaMemcpyDeviceToHost, st));
    TSL_CUDA(cudaStreamSynchronize(st));
    std::vector<uint32_t> sl(n);
    TSL_CUDA(cudaMemcpyAsync(sl.data(), o->sel.p + ((size_t)img * o->nlevels + level) * o->sel_cap, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    TSL_CUDA(cudaStreamSynchronize(st));
    if (out_bytes < (int)sizeof(int) * (1 + 3 * n)) return set_error(TSLAM_ERR_ARG, "buffer too small");
    io[0] = n;
    for (int q = 0; q < n; ++q) { io[1 + 3 * q] = (int)(sl[q] >> 18); io[2 + 3 * q] = (int)((sl[q] >> 8) & 1023u); io[3 + 3 * q] = (int)(sl[q] & 255u); }
    return TSLAM_OK;
  }
  return set_error(TSLAM_ERR_ARG, "bad what");
}

int tslam_orb_dev_bench(tslam_orb* o, const uint8_t* const* imgs, int n_imgs, int w, int hgt, int stride, int reps, float* ms_mean, int64_t* n_kp) {
  if (!o || !imgs || reps <= 0) return set_error(TSLAM_ERR_ARG, "bad argument");
  TSL_CUDA(cudaSetDevice(o->ctx->device));
  int rc = orb_configure(o, w, hgt, n_imgs);
  if (rc) return rc;
  if ((rc = orb_upload(o, imgs, n_imgs, w, hgt, stride))) return rc;   // inputs resident in HBM before the timed region
  cudaStream_t st = o->ctx->stream;
  if ((rc = orb_run(o, n_imgs))) return rc;                            // warm-up
  double total = 0;
  for (int r = 0; r < reps; ++r) {
    if ((rc = flush_l2(o->ctx))) return rc;
    TSL_CUDA(cudaEventRecord(o->ctx->ev0, st));
    if ((rc = orb_run(o, n_imgs))) return rc;
    TSL_CUDA(cudaEventRecord(o->ctx->ev1, st));
    TSL_CUDA(cudaEventSynchronize(o->ctx->ev1));
    float ms = 0;
    TSL_CUDA(cudaEventElapsedTime(&ms, o->ctx->ev0, o->ctx->ev1));
    total += ms;
  }
  std::vector<int> cnt(n_imgs);
  TSL_CUDA(cudaMemcpyAsync(cnt.data(), o->counts.p, sizeof(int) * n_imgs, cudaMemcpyDeviceToHost, st));
  if ((rc = orb_check_err(o))) return rc;
  long long s = 0; for (int c : cnt) s += c;
  if (ms_mean) *ms_mean = (float)(total / reps);
  if (n_kp) *n_kp = s;
  return TSLAM_OK;
}

}  // extern "C"
