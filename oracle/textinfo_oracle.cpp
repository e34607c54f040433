// ORACLE — TEST INFRASTRUCTURE ONLY. CPU restatement of tool::CalTextinfo + tool::CalStatistics
// (/root/reference/src/tool.cc:1178-1262): mean / standard deviation (n-1) of the image intensities inside the
// projected text quad, where "inside" is what cv::fillPoly paints (cv::Point truncation of the double vertices,
// 8-connected outline + scan-line interior).
//
// cv::fillPoly is OpenCV code absent from /root/reference; its rasteriser is restated here from its published
// algorithm (cv::clipLine + cv::LineIterator outline, 16.16 fixed-point edge walk with dx = trunc(dX/dY),
// spans ceil(x_left) .. floor(x_right), rows y0 <= y < y1) and pinned against Python cv2 4.13:
//   * quads entirely inside the image: bit-identical masks (tests/test_oracle_orb.py, golden + live, 20k random quads);
//   * quads crossing the image border: identical in ~99 % of random quads; cv2 >= 4.5 additionally paints isolated
//     border-column pixels for edges that leave the image (a clipped-endpoint refinement not reproduced here).
//     KNOWN DEVIATION, documented in DESIGN.md; OpenCV 3.3.1 (the README's version) predates that refinement.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace tso_ti {
typedef long long i64;
const int XS = 16; const i64 ONE = 1LL << 16;

static bool clip_line(int w, int h, i64& x1, i64& y1, i64& x2, i64& y2) {   // cv::clipLine
  const i64 right = w - 1, bottom = h - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    i64 a;
    if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (i64)((double)(a - y1) * (x2 - x1) / (y2 - y1)); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
    if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (i64)((double)(a - y2) * (x2 - x1) / (y2 - y1)); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) { a = c1 == 1 ? 0 : right; y1 += (i64)((double)(a - x1) * (y2 - y1) / (x2 - x1)); x1 = a; c1 = 0; }
      if (c2) { a = c2 == 1 ? 0 : right; y2 += (i64)((double)(a - x2) * (y2 - y1) / (x2 - x1)); x2 = a; c2 = 0; }
    }
  }
  return (c1 | c2) == 0;
}

// cv::Line (thickness 1, LINE_8): LineIterator(img, pt1, pt2, 8, leftToRight = true)
template <typename F> static void line_pixels(i64 x1, i64 y1, i64 x2, i64 y2, F put) {
  i64 dx = x2 - x1, dy = y2 - y1;
  int bx = 1, by = 1;
  if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }
  if (dy < 0) { dy = -dy; by = -1; }
  const bool swap = dy > dx;
  if (swap) std::swap(dx, dy);
  i64 err = dx - (dy + dy); const i64 plus = dx + dx, minus = -(dy + dy);
  i64 x = x1, y = y1;
  for (i64 i = 0; i <= dx; ++i) {
    put((int)x, (int)y);
    const bool m = err < 0;
    err += minus + (m ? plus : 0);
    if (swap) { y += by; if (m) x += bx; } else { x += bx; if (m) y += by; }
  }
}

static void fill_poly_mask(const int px[4], const int py[4], int w, int h, std::vector<uint8_t>& mask) {
  mask.assign((size_t)w * h, 0);
  struct Edge { int y0, y1; i64 x, dx; };
  std::vector<Edge> edges;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 3) & 3;  // pt0 = v[count-1] first
    i64 ax = px[j], ay = py[j], bx = px[i], by = py[i];
    i64 cx1 = ax, cy1 = ay, cx2 = bx, cy2 = by;
    const bool visible = clip_line(w, h, cx1, cy1, cx2, cy2);   // (cx, cy) are modified even when the result is "invisible", like cv::clipLine
    if (visible) line_pixels(cx1, cy1, cx2, cy2, [&](int x, int y) { mask[(size_t)y * w + x] = 1; });
    if (ay == by) continue;
    // edge for the scan-line fill: OpenCV >= 4.5 builds it from the CLIPPED end points when an end point lies outside
    // the image ("use clipped endpoints to create a more accurate PolyEdge"), from the vertices otherwise
    const bool outside = ax < 0 || ax >= w || bx < 0 || bx >= w || ay < 0 || ay >= h || by < 0 || by >= h;
    i64 e0x = ax << XS, e0y = ay, e1x = bx << XS, e1y = by;
    if (outside && cy1 != cy2) { e0x = cx1 << XS; e0y = cy1; e1x = cx2 << XS; e1y = cy2; }
    Edge e;
    e.dx = (e1x - e0x) / (e1y - e0y);   // C++ integer division: truncation toward zero
    if (ay < by) { e.y0 = (int)ay; e.y1 = (int)by; e.x = e0x + (ay - e0y) * e.dx; }
    else { e.y0 = (int)by; e.y1 = (int)ay; e.x = e1x + (by - e1y) * e.dx; }
    edges.push_back(e);
  }
  if (edges.empty()) return;
  int ymin = edges[0].y0, ymax = edges[0].y1;
  for (const Edge& e : edges) { ymin = std::min(ymin, e.y0); ymax = std::max(ymax, e.y1); }
  ymax = std::min(ymax, h);
  for (int y = std::max(ymin, 0); y < ymax; ++y) {
    i64 xs[4]; int n = 0;
    for (const Edge& e : edges) if (e.y0 <= y && y < e.y1) xs[n++] = e.x + (i64)(y - e.y0) * e.dx;
    std::sort(xs, xs + n);
    for (int k = 0; k + 1 < n; k += 2) {
      i64 x1 = (xs[k] + ONE - 1) >> XS, x2 = xs[k + 1] >> XS;
      if (x1 < w && x2 >= 0) { x1 = std::max<i64>(x1, 0); x2 = std::min<i64>(x2, w - 1); for (i64 x = x1; x <= x2; ++x) mask[(size_t)y * w + x] = 1; }
    }
  }
}
}  // namespace tso_ti

// quad: x0 y0 x1 y1 x2 y2 x3 y3 (doubles, image pixels). Returns 1 if (mu, sigma) are valid (tool.cc:1246-1262).
extern "C" int tso_text_info(const uint8_t* img, int w, int h, const double* quad, double* mu, double* sigma, uint8_t* mask_out) {
  using namespace tso_ti;
  int px[4], py[4];
  int xMin = w + 1, xMax = -1, yMin = h + 1, yMax = -1;          // tool.cc:1182-1194
  for (int i = 0; i < 4; ++i) {
    const double vx = quad[2 * i], vy = quad[2 * i + 1];
    px[i] = (int)vx; py[i] = (int)vy;                             // cv::Point(double, double): truncation
    if (vx > xMax) xMax = (int)std::ceil(vx);
    if (vx < xMin) xMin = (int)std::floor(vx);
    if (vy > yMax) yMax = (int)std::ceil(vy);
    if (vy < yMin) yMin = (int)std::floor(vy);
  }
  if (xMin < 0) xMin = 0; if (xMin >= w) xMin = w - 1;            // :1196-1211
  if (yMin < 0) yMin = 0; if (yMin >= h) yMin = h - 1;
  if (xMax >= w) xMax = w - 1; if (xMax < 0) xMax = 0;
  if (yMax >= h) yMax = h - 1; if (yMax < 0) yMax = 0;
  std::vector<uint8_t> mask;
  fill_poly_mask(px, py, w, h, mask);
  if (mask_out) std::memcpy(mask_out, mask.data(), mask.size());
  std::vector<double> v;
  for (int y = yMin; y <= yMax; ++y)
    for (int x = xMin; x <= xMax; ++x)
      if (mask[(size_t)y * w + x]) v.push_back((double)img[(size_t)y * w + x]);
  if (v.empty()) return 0;
  double s = 0; for (double t : v) s += t;
  *mu = s / (double)v.size();
  double ss = 0; for (double t : v) ss += (t - *mu) * (t - *mu);
  *sigma = std::sqrt(ss / (double)(v.size() - 1));
  return *sigma != 0 ? 1 : 0;
}
