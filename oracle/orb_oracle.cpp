// ORACLE — TEST INFRASTRUCTURE ONLY (see ba_math.hpp). CPU restatement of TextSLAM's ORB extractor
// (/root/reference/src/ORBextractor.cc, ORB-SLAM2 lineage) including the OpenCV primitives it calls.
//
// Pinning: the reference has no tests for this path. The OpenCV primitives restated here (FAST-9/16
// with non-max suppression, INTER_LINEAR u8 resize, 7x7 sigma=2 GaussianBlur, fastAtan2, cvRound) are
// checked bit-for-bit against Python cv2 4.13 in tests/test_oracle_orb.py and against committed golden
// vectors produced by cv2 (tests/golden/, script tests/golden/make_orb_golden.py); the orchestration
// above them (cells, quad-tree, orientation, steered BRIEF) is pinned against a second, independent
// restatement in plain Python on the real cv2 primitives (tests/test_oracle_orb_py.py): candidates,
// quad-tree winners, output order, angles and every descriptor byte agree exactly.
//
// Deliberate, documented deviations (SURVEY §7 hard part 2, Appendix C):
//  * DistributeOctTree sorts (size, node*) pairs (ORBextractor.cc:685): ties depend on heap addresses.
//    Here ties are broken by node creation order (later-created node first), which is what a fresh
//    heap gives; the GPU implementation uses the same rule.
//  * cos/sin of the keypoint angle (ORBextractor.cc:112-113) are evaluated by a fixed double-precision
//    operation sequence (textslam_b200/csrc/orb_math.h, shared with the GPU) and rounded to float
//    instead of calling the host libm's cosf/sinf.
//  * GaussianBlur taps: variant 0 = OpenCV >= 3.4 / 4.x fixed-point taps {18,34,48,56,48,34,18}/256
//    (the only variant checkable here), variant 1 = OpenCV 3.3.1's round(k*256) taps {18,34,49,55,49,34,18}
//    as recalled in SURVEY Appendix C (unverifiable here).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <vector>
#include <list>
#include <algorithm>
#include "../include/tslam_b200.h"
#include "../textslam_b200/csrc/orb_math.h"

namespace tso {

static const int8_t kPattern[1024] = {
#include "../textslam_b200/csrc/orb_pattern.inc"
};
const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19;

static inline int cv_round(double v) { return (int)std::nearbyint(v); }  // round-half-to-even (default FE mode)

struct Img {
  int w = 0, h = 0;
  std::vector<uint8_t> d;
  Img() {}
  Img(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_) {}
  const uint8_t* row(int y) const { return &d[(size_t)y * w]; }
  uint8_t* row(int y) { return &d[(size_t)y * w]; }
};

// ---- cv::resize INTER_LINEAR, 8UC1 (11-bit fixed point, two passes) -----------------
static void resize_coeffs(int sn, int dn, std::vector<int>& ofs, std::vector<short>& a0, std::vector<short>& a1) {
  const double scale = (double)sn / dn;
  ofs.resize(dn); a0.resize(dn); a1.resize(dn);
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    ofs[d] = s;
    a0[d] = (short)cv_round((1.f - f) * 2048.f);
    a1[d] = (short)cv_round(f * 2048.f);
  }
}
static void resize_linear(const Img& src, Img& dst) {
  std::vector<int> xo, yo; std::vector<short> xa0, xa1, ya0, ya1;
  resize_coeffs(src.w, dst.w, xo, xa0, xa1);
  resize_coeffs(src.h, dst.h, yo, ya0, ya1);
  std::vector<int> r0(dst.w), r1(dst.w);
  for (int y = 0; y < dst.h; ++y) {
    const int sy0 = yo[y], sy1 = std::min(sy0 + 1, src.h - 1);
    const uint8_t* p0 = src.row(sy0); const uint8_t* p1 = src.row(sy1);
    for (int x = 0; x < dst.w; ++x) {
      const int sx0 = xo[x], sx1 = std::min(sx0 + 1, src.w - 1);
      r0[x] = xa0[x] * p0[sx0] + xa1[x] * p0[sx1];
      r1[x] = xa0[x] * p1[sx0] + xa1[x] * p1[sx1];
    }
    uint8_t* o = dst.row(y);
    const int b0 = ya0[y], b1 = ya1[y];
    for (int x = 0; x < dst.w; ++x) o[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

// ---- cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101), 8UC1 -----------------------------------------
static inline int reflect101(int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; return p; }
static void gaussian7(const Img& src, Img& dst, int variant) {
  static const int T0[7] = {18, 34, 48, 56, 48, 34, 18};
  static const int T1[7] = {18, 34, 49, 55, 49, 34, 18};
  const int* T = variant == 1 ? T1 : T0;
  const int w = src.w, h = src.h;
  std::vector<int> tmp((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* p = src.row(y);
    for (int x = 0; x < w; ++x) { int s = 0; for (int k = 0; k < 7; ++k) s += T[k] * p[reflect101(x + k - 3, w)]; tmp[(size_t)y * w + x] = s; }
  }
  for (int y = 0; y < h; ++y) {
    uint8_t* o = dst.row(y);
    for (int x = 0; x < w; ++x) {
      int s = 0; for (int k = 0; k < 7; ++k) s += T[k] * tmp[(size_t)reflect101(y + k - 3, h) * w + x];
      const int v = (s + (1 << 15)) >> 16;  // FixedPtCastEx<int,uchar>(16): both variants differ only in the taps
      o[x] = (uint8_t)std::min(255, std::max(0, v));
    }
  }
}

// ---- cv::fastAtan2 (degrees) -----------------------------------------------------------------------------
static float fast_atan2(float y, float x) {
  const float k = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k, p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
  const float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) { c = ay / (ax + (float)DBL_EPSILON); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  else { c = ax / (ay + (float)DBL_EPSILON); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---- cv::FAST TYPE_9_16 corner measure: m = max over the 16 arcs of 9 of min(c_k - p) and min(p - c_k) --------------
static const int kCircle[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};
static inline int fast_measure(const uint8_t* p, int stride) {
  int d[16];
  const int v = p[0];
  for (int k = 0; k < 16; ++k) d[k] = (int)p[kCircle[k][1] * stride + kCircle[k][0]] - v;
  int best = -255;
  for (int s = 0; s < 16; ++s) {
    int mn = 255, mx = -255;
    for (int k = 0; k < 9; ++k) { const int t = d[(s + k) & 15]; mn = std::min(mn, t); mx = std::max(mx, t); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;  // corner at threshold t  <=>  best > t ; cornerScore = best - 1
}
struct RawKp { int x, y, resp; };
// cv::FAST(roi, kps, t, nonmaxSuppression=true) on the ROI [x0,x1) x [y0,y1): row-major output, ROI-relative coords.
static void fast_roi(const Img& im, int x0, int y0, int x1, int y1, int t, std::vector<RawKp>& out) {
  const int W = x1 - x0, H = y1 - y0;
  out.clear();
  if (W < 7 || H < 7) return;
  std::vector<int> score((size_t)W * H, 0);
  for (int y = 3; y < H - 3; ++y)
    for (int x = 3; x < W - 3; ++x) {
      const int m = fast_measure(im.row(y0 + y) + x0 + x, im.w);
      if (m > t) score[(size_t)y * W + x] = m - 1;
    }
  for (int y = 3; y < H - 3; ++y)
    for (int x = 3; x < W - 3; ++x) {
      const int s = score[(size_t)y * W + x];
      if (s == 0) continue;  // not a corner at this threshold (corners have score = m - 1 >= t >= 7 here)
      const int* c = &score[(size_t)y * W + x];
      if (s > c[-1] && s > c[1] && s > c[-W - 1] && s > c[-W] && s > c[-W + 1] && s > c[W - 1] && s > c[W] && s > c[W + 1])
        out.push_back(RawKp{x, y, s});
    }
}

// ---- quad-tree distribution (ORBextractor.cc:482-764) ---------------------------------------------------------
struct DKey { float x, y; int resp; };
struct Node {
  int ULx, ULy, URx, BLy;     // UL, UR.x, BL.y (BR = (URx, BLy))
  std::vector<int> keys;      // indices into the key array, in insertion order
  bool noMore = false;
  long seq = 0;               // creation order (list order == descending seq for pushed-front nodes)
  std::list<Node>::iterator lit;
};
static void divide(const Node& n, const std::vector<DKey>& K, Node c[4]) {
  const int halfX = (int)std::ceil((float)(n.URx - n.ULx) / 2), halfY = (int)std::ceil((float)(n.BLy - n.ULy) / 2);
  const int mx = n.ULx + halfX, my = n.ULy + halfY;
  c[0].ULx = n.ULx; c[0].ULy = n.ULy; c[0].URx = mx; c[0].BLy = my;
  c[1].ULx = mx; c[1].ULy = n.ULy; c[1].URx = n.URx; c[1].BLy = my;
  c[2].ULx = n.ULx; c[2].ULy = my; c[2].URx = mx; c[2].BLy = n.BLy;
  c[3].ULx = mx; c[3].ULy = my; c[3].URx = n.URx; c[3].BLy = n.BLy;
  for (int i : n.keys) {
    const DKey& k = K[i];
    if (k.x < mx) { if (k.y < my) c[0].keys.push_back(i); else c[2].keys.push_back(i); }
    else if (k.y < my) c[1].keys.push_back(i);
    else c[3].keys.push_back(i);
  }
  for (int q = 0; q < 4; ++q) c[q].noMore = c[q].keys.size() == 1;
}
static std::vector<int> distribute_octtree(const std::vector<DKey>& K, int minX, int maxX, int minY, int maxY, int N) {
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  const float hX = (float)(maxX - minX) / nIni;
  std::list<Node> L;
  std::vector<Node*> ini(nIni);
  long seq = 0;
  for (int i = 0; i < nIni; ++i) {
    Node n; n.ULx = (int)(hX * (float)i); n.ULy = 0; n.URx = (int)(hX * (float)(i + 1)); n.BLy = maxY - minY;
    n.seq = -(long)i;  // pushed back: list order ascending i == descending seq
    L.push_back(n); ini[i] = &L.back();
  }
  for (size_t i = 0; i < K.size(); ++i) ini[(int)(K[i].x / hX)]->keys.push_back((int)i);
  for (auto it = L.begin(); it != L.end();) {
    if (it->keys.size() == 1) { it->noMore = true; ++it; }
    else if (it->keys.empty()) it = L.erase(it);
    else ++it;
  }
  bool finish = false;
  std::vector<std::pair<int, Node*>> vSize;
  auto push_children = [&](Node c[4], int* nToExpand) {
    for (int q = 0; q < 4; ++q) {
      if (c[q].keys.empty()) continue;
      c[q].seq = ++seq;
      L.push_front(c[q]);
      if (c[q].keys.size() > 1) {
        if (nToExpand) ++*nToExpand;
        vSize.push_back(std::make_pair((int)c[q].keys.size(), &L.front()));
        L.front().lit = L.begin();
      }
    }
  };
  while (!finish) {
    const int prevSize = (int)L.size();
    int nToExpand = 0;
    vSize.clear();
    for (auto it = L.begin(); it != L.end();) {
      if (it->noMore) { ++it; continue; }
      Node c[4];
      divide(*it, K, c);
      push_children(c, &nToExpand);
      it = L.erase(it);
    }
    if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;
    else if ((int)L.size() + nToExpand * 3 > N) {
      while (!finish) {
        const int prev2 = (int)L.size();
        std::vector<std::pair<int, Node*>> vPrev = vSize;
        vSize.clear();
        // (size, pointer) ascending in the reference; pointer ties -> creation order (see header)
        std::sort(vPrev.begin(), vPrev.end(), [](const std::pair<int, Node*>& a, const std::pair<int, Node*>& b) {
          return a.first != b.first ? a.first < b.first : a.second->seq < b.second->seq;
        });
        for (int j = (int)vPrev.size() - 1; j >= 0; --j) {
          Node c[4];
          divide(*vPrev[j].second, K, c);
          push_children(c, nullptr);
          L.erase(vPrev[j].second->lit);
          if ((int)L.size() >= N) break;
        }
        if ((int)L.size() >= N || (int)L.size() == prev2) finish = true;
      }
    }
  }
  std::vector<int> res;
  for (auto& n : L) {
    int best = n.keys[0];
    for (size_t k = 1; k < n.keys.size(); ++k) if (K[n.keys[k]].resp > K[best].resp) best = n.keys[k];
    res.push_back(best);
  }
  return res;
}

struct Extractor {
  int nfeatures, nlevels, iniTh, minTh, blur_variant;
  float scaleFactor;
  std::vector<float> mvScale, mvInvScale;
  std::vector<int> perLevel, umax;
  std::vector<Img> pyr;
  Extractor(int nf, float sf, int nl, int ini, int mn, int bv) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), blur_variant(bv), scaleFactor(sf) {
    mvScale.resize(nl); mvInvScale.resize(nl);
    mvScale[0] = 1.0f;
    for (int i = 1; i < nl; ++i) mvScale[i] = mvScale[i - 1] * sf;
    for (int i = 0; i < nl; ++i) mvInvScale[i] = 1.0f / mvScale[i];
    perLevel.resize(nl);
    const float factor = 1.0f / sf;
    float nDesired = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) { perLevel[l] = cv_round(nDesired); sum += perLevel[l]; nDesired *= factor; }
    perLevel[nl - 1] = std::max(nf - sum, 0);
    umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) { while (umax[v0] == umax[v0 + 1]) ++v0; umax[v] = v0; ++v0; }
  }
  void compute_pyramid(const uint8_t* img, int w, int h, int stride) {  // ORBextractor.cc:1118-1143 (border never read, not kept)
    pyr.assign(nlevels, Img());
    for (int l = 0; l < nlevels; ++l) {
      const float s = mvInvScale[l];
      const int lw = cv_round((float)w * s), lh = cv_round((float)h * s);
      pyr[l] = Img(lw, lh);
      if (l == 0) for (int y = 0; y < h; ++y) std::memcpy(pyr[0].row(y), img + (size_t)y * stride, w);
      else resize_linear(pyr[l - 1], pyr[l]);
    }
  }
  float ic_angle(const Img& im, float px, float py) const {  // ORBextractor.cc:77-104
    int m01 = 0, m10 = 0;
    const int cx = cv_round(px), cy = cv_round(py);
    const uint8_t* c = im.row(cy) + cx;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m10 += u * c[u];
    const int step = im.w;
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
      int vs = 0; const int d = umax[v];
      for (int u = -d; u <= d; ++u) { const int vp = c[u + v * step], vm = c[u - v * step]; vs += vp - vm; m10 += u * (vp + vm); }
      m01 += v * vs;
    }
    return fast_atan2((float)m01, (float)m10);
  }
  void descriptor(const Img& im, float px, float py, float angle_deg, uint8_t* desc) const {  // ORBextractor.cc:108-147
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float angle = angle_deg * factorPI;
    double sd, cd;
    tsl_det_sincos((double)angle, &sd, &cd);   // see orb_math.h (deviation from the host-libm cosf/sinf, documented above)
    const float a = (float)cd, b = (float)sd;
    const uint8_t* center = im.row(cv_round(py)) + cv_round(px);
    const int step = im.w;
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
      int val = 0;
      for (int k = 0; k < 8; ++k) {
        const int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
        const int t0 = center[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
        const int t1 = center[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }
  std::vector<std::vector<DKey>> dbg_cand;
  std::vector<std::vector<int>> dbg_sel;
  // ORBextractor.cc:766-854 + :1054-1116
  int extract(const uint8_t* img, int w, int h, int stride, int max_kp, tslam_keypoint* kp_out, uint8_t* desc_out) {
    compute_pyramid(img, w, h, stride);
    int total = 0;
    const float W = 30;
    dbg_cand.assign(nlevels, {}); dbg_sel.assign(nlevels, {});
    for (int level = 0; level < nlevels; ++level) {
      const Img& im = pyr[level];
      const int minBX = EDGE_THRESHOLD - 3, minBY = minBX, maxBX = im.w - EDGE_THRESHOLD + 3, maxBY = im.h - EDGE_THRESHOLD + 3;
      std::vector<DKey> cand;
      const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
      const int nCols = (int)(width / W), nRows = (int)(height / W);
      const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
      std::vector<RawKp> cell;
      for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell); float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
          const float iniX = (float)(minBX + j * wCell); float maxX = iniX + wCell + 6;
          if (iniX >= maxBX - 6) continue;
          if (maxX > maxBX) maxX = (float)maxBX;
          fast_roi(im, (int)iniX, (int)iniY, (int)maxX, (int)maxY, iniTh, cell);
          if (cell.empty()) fast_roi(im, (int)iniX, (int)iniY, (int)maxX, (int)maxY, minTh, cell);
          for (const RawKp& k : cell) cand.push_back(DKey{(float)(k.x + j * wCell), (float)(k.y + i * hCell), k.resp});
        }
      }
      std::vector<int> sel;
      if (!cand.empty()) sel = distribute_octtree(cand, minBX, maxBX, minBY, maxBY, perLevel[level]);
      const int scaledPatch = (int)(PATCH_SIZE * mvScale[level]);
      dbg_cand[level] = cand; dbg_sel[level] = sel;
      if (sel.empty()) continue;
      Img blurred(im.w, im.h);
      gaussian7(im, blurred, blur_variant);
      for (int idx : sel) {
        if (total >= max_kp) return -1;
        float x = cand[idx].x + minBX, y = cand[idx].y + minBY;
        const float ang = ic_angle(im, x, y);
        descriptor(blurred, x, y, ang, desc_out + (size_t)total * 32);
        if (level != 0) { const float sc = mvScale[level]; x *= sc; y *= sc; }
        tslam_keypoint& o = kp_out[total];
        o.x = x; o.y = y; o.size = (float)scaledPatch; o.angle = ang; o.response = (float)cand[idx].resp; o.octave = level; o.class_id = -1;
        ++total;
      }
    }
    return total;
  }
};

}  // namespace tso

using namespace tso;

extern "C" {
int tso_orb_extract(const uint8_t* img, int w, int h, int stride, int nfeatures, float scale, int nlevels, int iniTh, int minTh, int blur_variant,
                    int max_kp, tslam_keypoint* kp, uint8_t* desc) {
  Extractor E(nfeatures, scale, nlevels, iniTh, minTh, blur_variant);
  return E.extract(img, w, h, stride, max_kp, kp, desc);
}
int tso_orb_level_size(int w, int h, float scale, int nlevels, int level, int* lw, int* lh) {
  Extractor E(1000, scale, nlevels, 20, 7, 0);
  *lw = cv_round((float)w * E.mvInvScale[level]); *lh = cv_round((float)h * E.mvInvScale[level]);
  return 0;
}
int tso_orb_pyramid_level(const uint8_t* img, int w, int h, int stride, float scale, int nlevels, int level, uint8_t* out) {
  Extractor E(1000, scale, nlevels, 20, 7, 0);
  E.compute_pyramid(img, w, h, stride);
  std::memcpy(out, E.pyr[level].d.data(), E.pyr[level].d.size());
  return 0;
}
int tso_orb_features_per_level(int nfeatures, float scale, int nlevels, int* out) {
  Extractor E(nfeatures, scale, nlevels, 20, 7, 0);
  for (int l = 0; l < nlevels; ++l) out[l] = E.perLevel[l];
  return 0;
}
void tso_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  Img s(sw, sh), d(dw, dh); std::memcpy(s.d.data(), src, s.d.size());
  resize_linear(s, d); std::memcpy(dst, d.d.data(), d.d.size());
}
void tso_gaussian7(const uint8_t* src, int w, int h, uint8_t* dst, int variant) {
  Img s(w, h), d(w, h); std::memcpy(s.d.data(), src, s.d.size());
  gaussian7(s, d, variant); std::memcpy(dst, d.d.data(), d.d.size());
}
int tso_fast(const uint8_t* img, int w, int h, int threshold, int max_kp, int* xyr) {
  Img s(w, h); std::memcpy(s.d.data(), img, s.d.size());
  std::vector<RawKp> out; fast_roi(s, 0, 0, w, h, threshold, out);
  const int n = std::min((int)out.size(), max_kp);
  for (int i = 0; i < n; ++i) { xyr[3 * i] = out[i].x; xyr[3 * i + 1] = out[i].y; xyr[3 * i + 2] = out[i].resp; }
  return (int)out.size();
}
float tso_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int tso_cv_round(double v) { return cv_round(v); }
}

extern "C" void tso_det_sincos(double x, double* s, double* c) { tsl_det_sincos(x, s, c); }

// stage-by-stage read-back for the GPU parity tests: what = 0 FAST measure plane (max(m,0), 0 where m <= min_th),
// 1 candidates of a level (vToDistributeKeys order), 2 quad-tree winners (list order). Returns the count.
extern "C" int tso_orb_debug(const uint8_t* img, int w, int h, int nfeatures, float scale, int nlevels, int iniTh, int minTh, int what, int level,
                             void* out, int max_items) {
  Extractor E(nfeatures, scale, nlevels, iniTh, minTh, 0);
  if (what == 0) {
    E.compute_pyramid(img, w, h, w);
    const Img& im = E.pyr[level];
    uint8_t* o = (uint8_t*)out;
    for (int y = 0; y < im.h; ++y) for (int x = 0; x < im.w; ++x) {
      int m = 0;
      if (x >= 3 && x < im.w - 3 && y >= 3 && y < im.h - 3) { m = fast_measure(im.row(y) + x, im.w); if (m <= minTh) m = 0; }
      o[(size_t)y * im.w + x] = (uint8_t)std::min(255, std::max(0, m));
    }
    return im.w * im.h;
  }
  std::vector<tslam_keypoint> kp(nfeatures + 4 * nlevels + 64); std::vector<uint8_t> desc(kp.size() * 32);
  E.extract(img, w, h, w, (int)kp.size(), kp.data(), desc.data());
  int* o = (int*)out;
  if (what == 1) {
    const auto& c = E.dbg_cand[level];
    const int n = std::min((int)c.size(), max_items);
    for (int i = 0; i < n; ++i) { o[3 * i] = (int)c[i].x; o[3 * i + 1] = (int)c[i].y; o[3 * i + 2] = c[i].resp; }
    return (int)c.size();
  }
  const auto& c = E.dbg_cand[level]; const auto& sl = E.dbg_sel[level];
  const int n = std::min((int)sl.size(), max_items);
  for (int i = 0; i < n; ++i) { o[3 * i] = (int)c[sl[i]].x; o[3 * i + 1] = (int)c[sl[i]].y; o[3 * i + 2] = c[sl[i]].resp; }
  return (int)sl.size();
}

// ---------------------------------------------------------------------------------------------------------
// Direct-method frame pyramid: frame::GetPyrMat (/root/reference/src/frame.cc:178-202) —
// cv::pyrDown chain, cv::Sobel(ddepth = CV_8U, ksize 3, BORDER_DEFAULT) in x and y, cv::addWeighted(0.5, 0.5).
// OpenCV rules pinned against cv2 in tests/test_oracle_orb.py: pyrDown = separable
// [1 4 6 4 1], REFLECT_101, (V + 128) >> 8, size ((w+1)/2, (h+1)/2); Sobel into CV_8U saturates (negative
// derivatives become 0); addWeighted rounds half to even.
// ---------------------------------------------------------------------------------------------------------
namespace tso {
static void pyr_down(const Img& src, Img& dst) {
  static const int K[5] = {1, 4, 6, 4, 1};
  for (int y = 0; y < dst.h; ++y)
    for (int x = 0; x < dst.w; ++x) {
      int s = 0;
      for (int j = 0; j < 5; ++j) {
        const uint8_t* row = src.row(reflect101(2 * y + j - 2, src.h));
        int rs = 0;
        for (int i = 0; i < 5; ++i) rs += K[i] * row[reflect101(2 * x + i - 2, src.w)];
        s += K[j] * rs;
      }
      dst.row(y)[x] = (uint8_t)((s + 128) >> 8);
    }
}
static void sobel_u8(const Img& src, Img& gx, Img& gy, Img& g) {
  for (int y = 0; y < src.h; ++y) {
    const uint8_t* r0 = src.row(reflect101(y - 1, src.h)); const uint8_t* r1 = src.row(y); const uint8_t* r2 = src.row(reflect101(y + 1, src.h));
    for (int x = 0; x < src.w; ++x) {
      const int xm = reflect101(x - 1, src.w), xp = reflect101(x + 1, src.w);
      const int dx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
      const int dy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
      const int a = std::min(255, std::max(0, dx)), b = std::min(255, std::max(0, dy));
      gx.row(y)[x] = (uint8_t)a; gy.row(y)[x] = (uint8_t)b;
      g.row(y)[x] = (uint8_t)std::min(255, cv_round(a * 0.5 + b * 0.5));
    }
  }
}
}  // namespace tso

// what: 0 image level, 1 grad (addWeighted), 2 grad_x, 3 grad_y. Returns level width * height.
extern "C" int tso_frame_pyramid(const uint8_t* img, int w, int h, int stride, int nlevels, int level, int what, uint8_t* out, int* lw, int* lh) {
  using namespace tso;
  Img cur(w, h);
  for (int y = 0; y < h; ++y) std::memcpy(cur.row(y), img + (size_t)y * stride, w);
  for (int l = 1; l <= level; ++l) { Img nxt((cur.w + 1) / 2, (cur.h + 1) / 2); pyr_down(cur, nxt); cur = nxt; }
  (void)nlevels;
  *lw = cur.w; *lh = cur.h;
  if (!out) return cur.w * cur.h;
  if (what == 0) { std::memcpy(out, cur.d.data(), cur.d.size()); return cur.w * cur.h; }
  Img gx(cur.w, cur.h), gy(cur.w, cur.h), g(cur.w, cur.h);
  sobel_u8(cur, gx, gy, g);
  const Img& o = what == 1 ? g : (what == 2 ? gx : gy);
  std::memcpy(out, o.d.data(), o.d.size());
  return cur.w * cur.h;
}
