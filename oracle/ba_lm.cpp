// ORACLE — TEST INFRASTRUCTURE ONLY (see ba_math.hpp header). PARITY UNPINNED against a real Ceres build; cross-checked
// instead against two independent re-implementations: a dense numpy reading of the same loop whose trace (accept / reject
// sequence, cost and radius of every iteration) it reproduces (tests/test_oracle_lm_py.py), and scipy.optimize.least_squares
// on the same robustified objective, which reaches the same optimum (tests/test_oracle_scipy.py).
//
// CPU restatement of what `ceres::Solve` does for the problems TextSLAM builds
// (/root/reference/src/optimizer.cc:1037-1044,1215-1222,1595-1602,1833-1840: TRUST_REGION +
// LEVENBERG_MARQUARDT, everything else default), followed by `Problem::Evaluate`
// (:1228-1233, :1609-1614). Ceres itself is NOT under /root/reference and not installed here;
// the loop below follows the published behaviour of Ceres 1.14-2.1 `TrustRegionMinimizer` /
// `LevenbergMarquardtStrategy` as summarised in SURVEY.md Appendix A:
//   Jacobi scaling 1/(1+||J_col||) fixed at iteration 0; D = sqrt(clamp(diag(J'J),1e-6,1e32)/radius);
//   exact solve of (J'J + D'D) y = J'r, step = -y; model_cost_change = -(Jd)'(r + Jd/2);
//   relative_decrease > 1e-3 accepts; radius /= max(1/3, 1-(2q-1)^3) on accept, radius /= k, k*=2 on
//   reject; parameter / function tolerance tested before the accept test; gradient tolerance and
//   max_num_iterations tested at the top of the loop.
// The linear system is solved exactly by eliminating the landmark blocks (rho 1x1, theta 3x3)
// and factoring the reduced camera matrix with an envelope Cholesky — the same step a sparse
// Cholesky of the full system yields (Appendix A.7). `dense_full=1` solves the full system
// with a dense Cholesky instead (cross-check for small problems).
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include "ba_math.hpp"
#include "../include/tslam_b200.h"

namespace tso {

struct Blocks {
  // per residual block: corrected residual and tangent Jacobian (scaled by Jacobi scaling later)
  std::vector<double> pr, pJ;  // n_pobs x 2, n_pobs x 26
  std::vector<double> tr, tJ;  // n_tobs x 8, n_tobs x 120
};

struct Layout {
  std::vector<int> camslot, lmslot, plslot;  // -1 = constant / unused
  int nc = 0, nl = 0, npl = 0;
  std::vector<uint8_t> p_active, t_active;   // block has at least one free parameter block
  int ncols() const { return 6 * nc + nl + 3 * npl; }
  int lmcol(int s) const { return 6 * nc + s; }
  int plcol(int s) const { return 6 * nc + nl + 3 * s; }
};

static Layout make_layout(const tslam_ba_problem& p) {
  Layout L;
  L.camslot.assign(p.n_cams, -1); L.lmslot.assign(p.n_points, -1); L.plslot.assign(p.n_planes, -1);
  std::vector<uint8_t> cu(p.n_cams, 0), lu(p.n_points, 0), pu(p.n_planes, 0);
  L.p_active.assign(p.n_pobs, 0); L.t_active.assign(p.n_tobs, 0);
  auto cf = [&](int k) { return p.cam_fixed && p.cam_fixed[k]; };
  for (int i = 0; i < p.n_pobs; ++i) {
    const int c = p.p_cam[i], h = p.p_host[i], l = p.p_lm[i];
    const bool lf = p.rho_fixed && p.rho_fixed[l];
    if (!cf(c) || !cf(h) || !lf) { L.p_active[i] = 1; cu[c] = cu[h] = 1; lu[l] = 1; }
  }
  for (int i = 0; i < p.n_tobs; ++i) {
    const int c = p.t_cam[i], h = p.t_host[i], l = p.t_plane[i];
    const bool lf = p.theta_fixed && p.theta_fixed[l];
    if (!cf(c) || !cf(h) || !lf) { L.t_active[i] = 1; cu[c] = cu[h] = 1; pu[l] = 1; }
  }
  for (int k = 0; k < p.n_cams; ++k) if (cu[k] && !cf(k)) L.camslot[k] = L.nc++;
  for (int k = 0; k < p.n_points; ++k) if (lu[k] && !(p.rho_fixed && p.rho_fixed[k])) L.lmslot[k] = L.nl++;
  for (int k = 0; k < p.n_planes; ++k) if (pu[k] && !(p.theta_fixed && p.theta_fixed[k])) L.plslot[k] = L.npl++;
  return L;
}

struct State {  // parameter values
  std::vector<double> cams, rho, theta;
};

static TextBlockConst text_const(const tslam_ba_problem& p, int i) {
  TextBlockConst c;
  c.img = p.imgs + (size_t)p.t_img[i] * p.img_w * p.img_h;
  c.cols = p.img_w; c.rows = p.img_h;
  c.rays = p.t_rays + 16 * (size_t)i; c.iref = p.t_iref + 8 * (size_t)i;
  c.mu = p.t_musigma[2 * i]; c.sigma = p.t_musigma[2 * i + 1];
  c.K4 = p.K_text; c.wT = p.w_text;
  return c;
}

// Evaluate all residual blocks at `x`. If B != null also Jacobians (loss-corrected). Returns the
// cost 0.5*sum rho(s) over ACTIVE blocks; fixed (all-constant) blocks go to *fixed_cost.
// If raw_out != null, the corrected residuals of every block (active or not) are written there.
static double evaluate(const tslam_ba_problem& p, const Layout& L, const State& x, int text_jac_mode, Blocks* B,
                       double* fixed_cost, double* raw_out, int n_threads) {
  const int NP = p.n_pobs, NT = p.n_tobs;
  if (B) { B->pr.resize(2 * (size_t)NP); B->pJ.resize(26 * (size_t)NP); B->tr.resize(8 * (size_t)NT); B->tJ.resize(120 * (size_t)NT); }
  std::vector<double> cost_p(NP), cost_t(NT);
  auto work_p = [&](int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      const double* cam = &x.cams[7 * p.p_cam[i]];
      const double* host = &x.cams[7 * p.p_host[i]];
      const double ray[3] = {p.p_ray[2 * i], p.p_ray[2 * i + 1], 1.0};
      double r[2], J[26];
      const bool wantJ = B && L.p_active[i];
      point_eval(cam, host, x.rho[p.p_lm[i]], ray, p.p_uv + 2 * i, p.K_point, p.w_point, r, wantJ ? J : nullptr);
      const double s = r[0] * r[0] + r[1] * r[1];
      double rho[3]; huber(p.huber_point, s, rho);
      const double sq = std::sqrt(rho[1]);
      cost_p[i] = 0.5 * rho[0];
      r[0] *= sq; r[1] *= sq;
      if (raw_out) { raw_out[2 * i] = r[0]; raw_out[2 * i + 1] = r[1]; }
      if (B) {
        B->pr[2 * (size_t)i] = r[0]; B->pr[2 * (size_t)i + 1] = r[1];
        double* Jo = &B->pJ[26 * (size_t)i];
        if (wantJ) for (int k = 0; k < 26; ++k) Jo[k] = J[k] * sq; else for (int k = 0; k < 26; ++k) Jo[k] = 0;
      }
    }
  };
  auto work_t = [&](int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      const double* cam = &x.cams[7 * p.t_cam[i]];
      const double* host = &x.cams[7 * p.t_host[i]];
      const double* th = &x.theta[3 * p.t_plane[i]];
      TextBlockConst c = text_const(p, i);
      double r[8], J[120];
      const bool wantJ = B && L.t_active[i];
      if (!wantJ) text_functor(cam, host, th, c, r);
      else if (text_jac_mode == TSLAM_JAC_CENTRAL_DIFF) {
        unsigned m = 0;
        if (L.camslot[p.t_cam[i]] >= 0) m |= 1u;
        if (L.camslot[p.t_host[i]] >= 0) m |= 2u;
        if (L.plslot[p.t_plane[i]] >= 0) m |= 4u;
        text_eval_numeric(cam, host, th, c, m, r, J);
      } else text_eval_analytic(cam, host, th, c, r, J);
      double s = 0; for (int k = 0; k < 8; ++k) s += r[k] * r[k];
      double rho[3]; huber(p.huber_text, s, rho);
      const double sq = std::sqrt(rho[1]);
      cost_t[i] = 0.5 * rho[0];
      for (int k = 0; k < 8; ++k) r[k] *= sq;
      if (raw_out) for (int k = 0; k < 8; ++k) raw_out[2 * (size_t)NP + 8 * (size_t)i + k] = r[k];
      if (B) {
        for (int k = 0; k < 8; ++k) B->tr[8 * (size_t)i + k] = r[k];
        double* Jo = &B->tJ[120 * (size_t)i];
        if (wantJ) for (int k = 0; k < 120; ++k) Jo[k] = J[k] * sq; else for (int k = 0; k < 120; ++k) Jo[k] = 0;
      }
    }
  };
  if (n_threads <= 1) { work_p(0, NP); work_t(0, NT); }
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
      th.emplace_back([&, t] { work_p((int)((long long)NP * t / n_threads), (int)((long long)NP * (t + 1) / n_threads));
                               work_t((int)((long long)NT * t / n_threads), (int)((long long)NT * (t + 1) / n_threads)); });
    for (auto& t : th) t.join();
  }
  double cost = 0, fixed = 0;
  for (int i = 0; i < NP; ++i) (L.p_active[i] ? cost : fixed) += cost_p[i];
  for (int i = 0; i < NT; ++i) (L.t_active[i] ? cost : fixed) += cost_t[i];
  if (fixed_cost) *fixed_cost = fixed;
  return cost;
}

// column index helpers: tangent column of (block kind, slot)
struct Cols { int c, h, l; };

// ------------------------------------------------------------------------------------------
// Envelope (skyline) Cholesky of a dense-stored symmetric matrix (lower triangle used).
// first[i] = first structurally non-zero column of row i. Returns false if not positive definite.
// ------------------------------------------------------------------------------------------
static bool envelope_cholesky(std::vector<double>& A, int n, const std::vector<int>& first) {
  for (int i = 0; i < n; ++i) {
    double* Ai = &A[(size_t)i * n];
    for (int j = first[i]; j <= i; ++j) {
      const double* Aj = &A[(size_t)j * n];
      const int k0 = std::max(first[i], first[j]);
      double s = Ai[j];
      for (int k = k0; k < j; ++k) s -= Ai[k] * Aj[k];
      if (j < i) Ai[j] = s / Aj[j];
      else { if (!(s > 0.0)) return false; Ai[i] = std::sqrt(s); }
    }
  }
  return true;
}
static void envelope_solve(const std::vector<double>& A, int n, const std::vector<int>& first, std::vector<double>& b) {
  for (int i = 0; i < n; ++i) {  // L y = b
    const double* Ai = &A[(size_t)i * n];
    double s = b[i];
    for (int k = first[i]; k < i; ++k) s -= Ai[k] * b[k];
    b[i] = s / Ai[i];
  }
  for (int i = n - 1; i >= 0; --i) {  // L' x = y (column sweep)
    const double* Ai = &A[(size_t)i * n];
    b[i] /= Ai[i];
    const double xi = b[i];
    for (int k = first[i]; k < i; ++k) b[k] -= Ai[k] * xi;
  }
}

struct Linear {  // scaled normal equations in Schur form
  int nc, nl, npl;
  std::vector<double> U;               // 6nc x 6nc (full symmetric storage)
  std::vector<double> gc;              // 6nc
  std::vector<double> V1, g1;          // nl (1x1 blocks)
  std::vector<double> V3, g3;          // npl x 9, npl x 3
  // cam-landmark blocks: per landmark a short list of (camslot, 6xD block)
  struct E1 { int cs; double e[6]; };
  struct E3 { int cs; double e[18]; };  // 6x3 row-major
  std::vector<std::vector<E1>> W1;
  std::vector<std::vector<E3>> W3;
};

static void accumulate(const tslam_ba_problem& p, const Layout& L, const Blocks& B, const std::vector<double>& scale, Linear& N) {
  const int nc = L.nc, n6 = 6 * nc;
  N.nc = nc; N.nl = L.nl; N.npl = L.npl;
  N.U.assign((size_t)n6 * n6, 0.0); N.gc.assign(n6, 0.0);
  N.V1.assign(L.nl, 0.0); N.g1.assign(L.nl, 0.0);
  N.V3.assign(9 * (size_t)L.npl, 0.0); N.g3.assign(3 * (size_t)L.npl, 0.0);
  N.W1.assign(L.nl, {}); N.W3.assign(L.npl, {});
  auto addU = [&](int sa, int sb, const double* Ja, const double* Jb, int ld, int rows) {
    // U[sa,sb] += Ja' Jb  (Ja,Jb: rows x 6 with leading dim ld)
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) {
      double s = 0; for (int k = 0; k < rows; ++k) s += Ja[k * ld + a] * Jb[k * ld + b];
      N.U[(size_t)(6 * sa + a) * n6 + 6 * sb + b] += s;
    }
  };
  for (int i = 0; i < p.n_pobs; ++i) {
    if (!L.p_active[i]) continue;
    const int cs = L.camslot[p.p_cam[i]], hs = L.camslot[p.p_host[i]], ls = L.lmslot[p.p_lm[i]];
    double J[26]; const double* Jr = &B.pJ[26 * (size_t)i]; const double* r = &B.pr[2 * (size_t)i];
    for (int k = 0; k < 2; ++k) for (int c = 0; c < 13; ++c) {
      double sc = 0.0;
      if (c < 6) sc = cs >= 0 ? scale[6 * cs + c] : 0.0;
      else if (c < 12) sc = hs >= 0 ? scale[6 * hs + c - 6] : 0.0;
      else sc = ls >= 0 ? scale[L.lmcol(ls)] : 0.0;
      J[13 * k + c] = Jr[13 * k + c] * sc;
    }
    if (cs >= 0) { addU(cs, cs, J, J, 13, 2); for (int a = 0; a < 6; ++a) N.gc[6 * cs + a] += J[a] * r[0] + J[13 + a] * r[1]; }
    if (hs >= 0) { addU(hs, hs, J + 6, J + 6, 13, 2); for (int a = 0; a < 6; ++a) N.gc[6 * hs + a] += J[6 + a] * r[0] + J[19 + a] * r[1]; }
    if (cs >= 0 && hs >= 0) {
      if (cs != hs) { addU(cs, hs, J, J + 6, 13, 2); addU(hs, cs, J + 6, J, 13, 2); }
      else { addU(cs, cs, J, J + 6, 13, 2); addU(cs, cs, J + 6, J, 13, 2); }
    }
    if (ls >= 0) {
      N.V1[ls] += J[12] * J[12] + J[25] * J[25];
      N.g1[ls] += J[12] * r[0] + J[25] * r[1];
      auto addW = [&](int s, const double* Jc) {
        auto& lst = N.W1[ls]; Linear::E1* e = nullptr;
        for (auto& q : lst) if (q.cs == s) { e = &q; break; }
        if (!e) { lst.push_back(Linear::E1{s, {0, 0, 0, 0, 0, 0}}); e = &lst.back(); }
        for (int a = 0; a < 6; ++a) e->e[a] += Jc[a] * J[12] + Jc[13 + a] * J[25];
      };
      if (cs >= 0) addW(cs, J);
      if (hs >= 0) addW(hs, J + 6);
    }
  }
  for (int i = 0; i < p.n_tobs; ++i) {
    if (!L.t_active[i]) continue;
    const int cs = L.camslot[p.t_cam[i]], hs = L.camslot[p.t_host[i]], ls = L.plslot[p.t_plane[i]];
    double J[120]; const double* Jr = &B.tJ[120 * (size_t)i]; const double* r = &B.tr[8 * (size_t)i];
    for (int k = 0; k < 8; ++k) for (int c = 0; c < 15; ++c) {
      double sc = 0.0;
      if (c < 6) sc = cs >= 0 ? scale[6 * cs + c] : 0.0;
      else if (c < 12) sc = hs >= 0 ? scale[6 * hs + c - 6] : 0.0;
      else sc = ls >= 0 ? scale[L.plcol(ls) + c - 12] : 0.0;
      J[15 * k + c] = Jr[15 * k + c] * sc;
    }
    auto gacc = [&](int s, const double* Jc) { for (int a = 0; a < 6; ++a) { double t = 0; for (int k = 0; k < 8; ++k) t += Jc[15 * k + a] * r[k]; N.gc[6 * s + a] += t; } };
    if (cs >= 0) { addU(cs, cs, J, J, 15, 8); gacc(cs, J); }
    if (hs >= 0) { addU(hs, hs, J + 6, J + 6, 15, 8); gacc(hs, J + 6); }
    if (cs >= 0 && hs >= 0) {
      if (cs != hs) { addU(cs, hs, J, J + 6, 15, 8); addU(hs, cs, J + 6, J, 15, 8); }
      else { addU(cs, cs, J, J + 6, 15, 8); addU(cs, cs, J + 6, J, 15, 8); }
    }
    if (ls >= 0) {
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) { double t = 0; for (int k = 0; k < 8; ++k) t += J[15 * k + 12 + a] * J[15 * k + 12 + b]; N.V3[9 * (size_t)ls + 3 * a + b] += t; }
        double t = 0; for (int k = 0; k < 8; ++k) t += J[15 * k + 12 + a] * r[k]; N.g3[3 * (size_t)ls + a] += t;
      }
      auto addW = [&](int s, const double* Jc) {
        auto& lst = N.W3[ls]; Linear::E3* e = nullptr;
        for (auto& q : lst) if (q.cs == s) { e = &q; break; }
        if (!e) { lst.push_back(Linear::E3{s, {}}); e = &lst.back(); for (double& v : e->e) v = 0; }
        for (int a = 0; a < 6; ++a) for (int b = 0; b < 3; ++b) { double t = 0; for (int k = 0; k < 8; ++k) t += Jc[15 * k + a] * J[15 * k + 12 + b]; e->e[3 * a + b] += t; }
      };
      if (cs >= 0) addW(cs, J);
      if (hs >= 0) addW(hs, J + 6);
    }
  }
}

static bool inv3_spd(const double* A, double* Ai) {
  const double a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[8];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (!(det > 0.0) || !(a > 0.0)) return false;
  const double id = 1.0 / det;
  Ai[0] = c00 * id; Ai[1] = c01 * id; Ai[2] = c02 * id;
  Ai[3] = Ai[1]; Ai[4] = (a * f - c * c) * id; Ai[5] = (b * c - a * e) * id;
  Ai[6] = Ai[2]; Ai[7] = Ai[5]; Ai[8] = (a * d - b * b) * id;
  return true;
}

// Solve (H + D^2) y = g for the scaled system; y laid out as [cams(6nc) | rho(nl) | theta(3npl)].
static bool solve_schur(const Linear& N, double radius, std::vector<double>& y) {
  const int nc = N.nc, n6 = 6 * nc;
  auto dmp = [&](double d) { return std::min(std::max(d, 1e-6), 1e32) / radius; };
  std::vector<double> S = N.U, b = N.gc;
  for (int i = 0; i < n6; ++i) S[(size_t)i * n6 + i] += dmp(N.U[(size_t)i * n6 + i]);
  std::vector<double> Vi1(N.nl), Vi3(9 * (size_t)N.npl);
  for (int l = 0; l < N.nl; ++l) {
    const double v = N.V1[l] + dmp(N.V1[l]);
    const double vi = 1.0 / v; Vi1[l] = vi;
    const auto& lst = N.W1[l];
    for (const auto& ea : lst) {
      for (int a = 0; a < 6; ++a) b[6 * ea.cs + a] -= ea.e[a] * vi * N.g1[l];
      for (const auto& eb : lst)
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c)
          S[(size_t)(6 * ea.cs + a) * n6 + 6 * eb.cs + c] -= ea.e[a] * vi * eb.e[c];
    }
  }
  for (int l = 0; l < N.npl; ++l) {
    double V[9]; for (int k = 0; k < 9; ++k) V[k] = N.V3[9 * (size_t)l + k];
    for (int a = 0; a < 3; ++a) V[4 * a] += dmp(N.V3[9 * (size_t)l + 4 * a]);
    double* Vi = &Vi3[9 * (size_t)l];
    if (!inv3_spd(V, Vi)) return false;
    const double* g = &N.g3[3 * (size_t)l];
    const auto& lst = N.W3[l];
    for (const auto& ea : lst) {
      double EV[18];  // 6x3 = E * Vi
      for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c) EV[3 * a + c] = ea.e[3 * a] * Vi[c] + ea.e[3 * a + 1] * Vi[3 + c] + ea.e[3 * a + 2] * Vi[6 + c];
      for (int a = 0; a < 6; ++a) b[6 * ea.cs + a] -= EV[3 * a] * g[0] + EV[3 * a + 1] * g[1] + EV[3 * a + 2] * g[2];
      for (const auto& eb : lst)
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c)
          S[(size_t)(6 * ea.cs + a) * n6 + 6 * eb.cs + c] -= EV[3 * a] * eb.e[3 * c] + EV[3 * a + 1] * eb.e[3 * c + 1] + EV[3 * a + 2] * eb.e[3 * c + 2];
    }
  }
  // envelope of the reduced matrix
  std::vector<int> first(n6);
  for (int i = 0; i < n6; ++i) { int f = i; const double* r = &S[(size_t)i * n6]; for (int j = 0; j < i; ++j) if (r[j] != 0.0) { f = j; break; } first[i] = f; }
  // block rows share an envelope start (keeps the 6x6 structure intact)
  for (int bi = 0; bi < nc; ++bi) { int f = n6; for (int a = 0; a < 6; ++a) f = std::min(f, first[6 * bi + a]); f = (f / 6) * 6; for (int a = 0; a < 6; ++a) first[6 * bi + a] = std::min(f, 6 * bi + a); }
  if (n6 > 0) {
    if (!envelope_cholesky(S, n6, first)) return false;
    envelope_solve(S, n6, first, b);
  }
  y.assign((size_t)n6 + N.nl + 3 * (size_t)N.npl, 0.0);
  for (int i = 0; i < n6; ++i) y[i] = b[i];
  for (int l = 0; l < N.nl; ++l) {
    double t = N.g1[l];
    for (const auto& e : N.W1[l]) for (int a = 0; a < 6; ++a) t -= e.e[a] * y[6 * e.cs + a];
    y[n6 + l] = t * Vi1[l];
  }
  for (int l = 0; l < N.npl; ++l) {
    double t[3] = {N.g3[3 * (size_t)l], N.g3[3 * (size_t)l + 1], N.g3[3 * (size_t)l + 2]};
    for (const auto& e : N.W3[l]) for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c) t[c] -= e.e[3 * a + c] * y[6 * e.cs + a];
    const double* Vi = &Vi3[9 * (size_t)l];
    for (int c = 0; c < 3; ++c) y[n6 + N.nl + 3 * l + c] = Vi[3 * c] * t[0] + Vi[3 * c + 1] * t[1] + Vi[3 * c + 2] * t[2];
  }
  for (double v : y) if (!std::isfinite(v)) return false;
  return true;
}

// Cross-check path: assemble the full dense (H + D^2) and factor it.
static bool solve_dense_full(const Linear& N, double radius, std::vector<double>& y) {
  const int n6 = 6 * N.nc, n = n6 + N.nl + 3 * N.npl;
  std::vector<double> H((size_t)n * n, 0.0), g(n, 0.0);
  for (int i = 0; i < n6; ++i) { g[i] = N.gc[i]; for (int j = 0; j < n6; ++j) H[(size_t)i * n + j] = N.U[(size_t)i * n6 + j]; }
  for (int l = 0; l < N.nl; ++l) {
    const int c = n6 + l; H[(size_t)c * n + c] = N.V1[l]; g[c] = N.g1[l];
    for (const auto& e : N.W1[l]) for (int a = 0; a < 6; ++a) { H[(size_t)(6 * e.cs + a) * n + c] = e.e[a]; H[(size_t)c * n + 6 * e.cs + a] = e.e[a]; }
  }
  for (int l = 0; l < N.npl; ++l) {
    const int c = n6 + N.nl + 3 * l;
    for (int a = 0; a < 3; ++a) { g[c + a] = N.g3[3 * (size_t)l + a]; for (int b2 = 0; b2 < 3; ++b2) H[(size_t)(c + a) * n + c + b2] = N.V3[9 * (size_t)l + 3 * a + b2]; }
    for (const auto& e : N.W3[l]) for (int a = 0; a < 6; ++a) for (int b2 = 0; b2 < 3; ++b2) { H[(size_t)(6 * e.cs + a) * n + c + b2] = e.e[3 * a + b2]; H[(size_t)(c + b2) * n + 6 * e.cs + a] = e.e[3 * a + b2]; }
  }
  for (int i = 0; i < n; ++i) H[(size_t)i * n + i] += std::min(std::max(H[(size_t)i * n + i], 1e-6), 1e32) / radius;
  std::vector<int> first(n, 0);
  if (!envelope_cholesky(H, n, first)) return false;
  envelope_solve(H, n, first, g);
  y = g;
  for (double v : y) if (!std::isfinite(v)) return false;
  return true;
}

static void apply_step(const tslam_ba_problem& p, const Layout& L, const State& x, const std::vector<double>& delta, State& out) {
  out = x;
  for (int k = 0; k < p.n_cams; ++k) {
    const int s = L.camslot[k]; if (s < 0) continue;
    quat_plus(&x.cams[7 * k], &delta[6 * s], &out.cams[7 * k]);
    for (int a = 0; a < 3; ++a) out.cams[7 * k + 4 + a] = x.cams[7 * k + 4 + a] + delta[6 * s + 3 + a];
  }
  for (int k = 0; k < p.n_points; ++k) { const int s = L.lmslot[k]; if (s >= 0) out.rho[k] = x.rho[k] + delta[L.lmcol(s)]; }
  for (int k = 0; k < p.n_planes; ++k) { const int s = L.plslot[k]; if (s >= 0) for (int a = 0; a < 3; ++a) out.theta[3 * k + a] = x.theta[3 * k + a] + delta[L.plcol(s) + a]; }
}
static double free_norm(const tslam_ba_problem& p, const Layout& L, const State& a, const State* b) {
  double s = 0;
  auto d = [&](const std::vector<double>& u, const std::vector<double>* v, size_t i) { const double t = v ? u[i] - (*v)[i] : u[i]; return t * t; };
  for (int k = 0; k < p.n_cams; ++k) if (L.camslot[k] >= 0) for (int c = 0; c < 7; ++c) s += d(a.cams, b ? &b->cams : nullptr, 7 * (size_t)k + c);
  for (int k = 0; k < p.n_points; ++k) if (L.lmslot[k] >= 0) s += d(a.rho, b ? &b->rho : nullptr, k);
  for (int k = 0; k < p.n_planes; ++k) if (L.plslot[k] >= 0) for (int c = 0; c < 3; ++c) s += d(a.theta, b ? &b->theta : nullptr, 3 * (size_t)k + c);
  return std::sqrt(s);
}

// unscaled tangent gradient and the Ceres gradient max norm ||x - Plus(x,-g)||_inf
static double gradient_max_norm(const tslam_ba_problem& p, const Layout& L, const State& x, const Linear& N, const std::vector<double>& scale) {
  // N holds the SCALED gradient g_s = S g; unscaled g = g_s / s
  double m = 0;
  for (int k = 0; k < p.n_cams; ++k) {
    const int s = L.camslot[k]; if (s < 0) continue;
    double ng[3]; for (int a = 0; a < 3; ++a) ng[a] = -N.gc[6 * s + a] / scale[6 * s + a];
    double q[4]; quat_plus(&x.cams[7 * k], ng, q);
    for (int a = 0; a < 4; ++a) m = std::max(m, std::fabs(x.cams[7 * k + a] - q[a]));
    for (int a = 3; a < 6; ++a) m = std::max(m, std::fabs(N.gc[6 * s + a] / scale[6 * s + a]));
  }
  for (int l = 0; l < N.nl; ++l) m = std::max(m, std::fabs(N.g1[l] / scale[L.lmcol(l)]));
  for (int l = 0; l < N.npl; ++l) for (int a = 0; a < 3; ++a) m = std::max(m, std::fabs(N.g3[3 * (size_t)l + a] / scale[L.plcol(l) + a]));
  return m;
}

static double model_cost_change(const tslam_ba_problem& p, const Layout& L, const Blocks& B, const std::vector<double>& delta) {
  // -(J d)'(r + J d / 2), J unscaled-corrected, d unscaled
  double acc = 0;
  for (int i = 0; i < p.n_pobs; ++i) {
    if (!L.p_active[i]) continue;
    const int cs = L.camslot[p.p_cam[i]], hs = L.camslot[p.p_host[i]], ls = L.lmslot[p.p_lm[i]];
    const double* J = &B.pJ[26 * (size_t)i]; const double* r = &B.pr[2 * (size_t)i];
    for (int k = 0; k < 2; ++k) {
      double m = 0;
      if (cs >= 0) for (int a = 0; a < 6; ++a) m += J[13 * k + a] * delta[6 * cs + a];
      if (hs >= 0) for (int a = 0; a < 6; ++a) m += J[13 * k + 6 + a] * delta[6 * hs + a];
      if (ls >= 0) m += J[13 * k + 12] * delta[L.lmcol(ls)];
      acc -= m * (r[k] + m / 2.0);
    }
  }
  for (int i = 0; i < p.n_tobs; ++i) {
    if (!L.t_active[i]) continue;
    const int cs = L.camslot[p.t_cam[i]], hs = L.camslot[p.t_host[i]], ls = L.plslot[p.t_plane[i]];
    const double* J = &B.tJ[120 * (size_t)i]; const double* r = &B.tr[8 * (size_t)i];
    for (int k = 0; k < 8; ++k) {
      double m = 0;
      if (cs >= 0) for (int a = 0; a < 6; ++a) m += J[15 * k + a] * delta[6 * cs + a];
      if (hs >= 0) for (int a = 0; a < 6; ++a) m += J[15 * k + 6 + a] * delta[6 * hs + a];
      if (ls >= 0) for (int a = 0; a < 3; ++a) m += J[15 * k + 12 + a] * delta[L.plcol(ls) + a];
      acc -= m * (r[k] + m / 2.0);
    }
  }
  return acc;
}

}  // namespace tso

using namespace tso;

extern "C" int tso_solve(tslam_ba_problem* p, const tslam_solve_options* opt, tslam_solve_summary* sum, double* final_residuals, double* trace) {
  auto T0 = std::chrono::steady_clock::now();
  const double ftol = opt->function_tolerance > 0 ? opt->function_tolerance : 1e-6;
  const double gtol = opt->gradient_tolerance > 0 ? opt->gradient_tolerance : 1e-10;
  const double ptol = opt->parameter_tolerance > 0 ? opt->parameter_tolerance : 1e-8;
  double radius = opt->initial_radius > 0 ? opt->initial_radius : 1e4;
  const double max_radius = 1e16, min_radius = 1e-32, min_rel_dec = 1e-3;
  const int nth = opt->n_threads > 1 ? opt->n_threads : 1;
  double decrease_factor = 2.0;

  Layout L = make_layout(*p);
  State x; x.cams.assign(p->cams, p->cams + 7 * (size_t)p->n_cams); x.rho.assign(p->rho, p->rho + p->n_points);
  x.theta.assign(p->theta, p->theta + 3 * (size_t)p->n_planes);
  Blocks B; Linear N;
  tslam_solve_summary S{}; S.n_free_cams = L.nc; S.n_free_points = L.nl; S.n_free_planes = L.npl; S.reduced_dim = 6 * L.nc;
  auto T1 = std::chrono::steady_clock::now();

  double fixed_cost = 0;
  double x_cost = evaluate(*p, L, x, opt->text_jac_mode, &B, &fixed_cost, nullptr, nth);
  S.initial_cost = x_cost + fixed_cost; S.fixed_cost = fixed_cost;
  const int ncols = L.ncols();
  // Jacobi scaling from the iteration-0 Jacobian
  std::vector<double> scale(ncols, 1.0);
  {
    std::vector<double> ones(ncols, 1.0);
    accumulate(*p, L, B, ones, N);
    const int n6 = 6 * L.nc;
    for (int i = 0; i < n6; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(N.U[(size_t)i * n6 + i]));
    for (int l = 0; l < L.nl; ++l) scale[L.lmcol(l)] = 1.0 / (1.0 + std::sqrt(N.V1[l]));
    for (int l = 0; l < L.npl; ++l) for (int a = 0; a < 3; ++a) scale[L.plcol(l) + a] = 1.0 / (1.0 + std::sqrt(N.V3[9 * (size_t)l + 4 * a]));
  }
  accumulate(*p, L, B, scale, N);
  double x_norm = free_norm(*p, L, x, nullptr);
  double gmax = gradient_max_norm(*p, L, x, N, scale);
  int iter = 0, n_ok = 0, n_bad = 0, term = TSLAM_TERM_NO_CONVERGENCE, invalid_run = 0;
  if (trace) { trace[0] = x_cost + fixed_cost; trace[1] = radius; trace[2] = 0; trace[3] = 1; }
  State cand;
  bool last_successful = true;
  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iter >= opt->max_iters) { term = TSLAM_TERM_NO_CONVERGENCE; break; }
    if (last_successful && gmax <= gtol) { term = TSLAM_TERM_GRADIENT_TOL; break; }
    if (radius <= min_radius) { term = TSLAM_TERM_NO_CONVERGENCE; break; }
    if (ncols == 0) break;
    ++iter;
    std::vector<double> y;
    const bool ok = opt->dense_full ? solve_dense_full(N, radius, y) : solve_schur(N, radius, y);
    double mcc = 0; std::vector<double> delta(ncols);
    if (ok) {
      for (int i = 0; i < ncols; ++i) delta[i] = -y[i] * scale[i];
      mcc = model_cost_change(*p, L, B, delta);
    }
    if (!ok || !(mcc > 0.0)) {  // invalid step
      ++invalid_run; ++n_bad; last_successful = false;
      if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = 0; t[3] = -1; }
      if (invalid_run >= 5) { term = TSLAM_TERM_FAILURE; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      continue;
    }
    invalid_run = 0;
    apply_step(*p, L, x, delta, cand);
    const double cand_cost = evaluate(*p, L, cand, opt->text_jac_mode, nullptr, nullptr, nullptr, nth);
    const double step_norm = free_norm(*p, L, x, &cand);
    if (step_norm <= ptol * (x_norm + ptol)) { term = TSLAM_TERM_PARAMETER_TOL; if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = 0; t[3] = 0; } break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= ftol * x_cost) { term = TSLAM_TERM_FUNCTION_TOL; if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = 0; t[3] = 0; } break; }
    const double rel = cost_change / mcc;
    if (rel > min_rel_dec) {
      x = cand; x_norm = free_norm(*p, L, x, nullptr);
      x_cost = evaluate(*p, L, x, opt->text_jac_mode, &B, nullptr, nullptr, nth);
      accumulate(*p, L, B, scale, N);
      gmax = gradient_max_norm(*p, L, x, N, scale);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(max_radius, radius); decrease_factor = 2.0;
      ++n_ok; last_successful = true;
      if (trace) { double* t = trace + 4 * iter; t[0] = x_cost + fixed_cost; t[1] = radius; t[2] = rel; t[3] = 1; }
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      ++n_bad; last_successful = false;
      if (trace) { double* t = trace + 4 * iter; t[0] = cand_cost + fixed_cost; t[1] = radius; t[2] = rel; t[3] = 0; }
    }
  }
  auto T2 = std::chrono::steady_clock::now();
  std::copy(x.cams.begin(), x.cams.end(), p->cams);
  std::copy(x.rho.begin(), x.rho.end(), p->rho);
  std::copy(x.theta.begin(), x.theta.end(), p->theta);
  if (final_residuals) evaluate(*p, L, x, opt->text_jac_mode, nullptr, nullptr, final_residuals, nth);
  auto T3 = std::chrono::steady_clock::now();
  S.iterations = iter; S.successful_steps = n_ok; S.unsuccessful_steps = n_bad; S.termination = term;
  S.final_cost = x_cost + fixed_cost;
  S.setup_ms = std::chrono::duration<double, std::milli>(T1 - T0).count();
  S.solve_ms = std::chrono::duration<double, std::milli>(T2 - T1).count();
  S.total_ms = std::chrono::duration<double, std::milli>(T3 - T0).count();
  if (sum) *sum = S;
  return 0;
}

// Residual + tangent Jacobian of every point block (kind = TSLAM_PT_*); mirrors tslam_eval_points.
extern "C" int tso_eval_points(int kind, const tslam_ba_problem* p, double* r, double* J, int n_threads) {
  const int nc = kind == TSLAM_PT_BA || kind == TSLAM_PT_BA_NW ? 13 : (kind == TSLAM_PT_POSE ? 6 : 1);
  const double w1[2] = {1.0, 1.0};
  const double* w = (kind == TSLAM_PT_BA_NW || kind == TSLAM_PT_RHO) ? w1 : p->w_point;
  auto work = [&](int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      const double ray[3] = {p->p_ray[2 * i], p->p_ray[2 * i + 1], 1.0};
      double rr[2], JJ[26];
      point_eval(p->cams + 7 * (size_t)p->p_cam[i], p->cams + 7 * (size_t)p->p_host[i], p->rho[p->p_lm[i]], ray, p->p_uv + 2 * i, p->K_point, w, rr, J ? JJ : nullptr);
      r[2 * (size_t)i] = rr[0]; r[2 * (size_t)i + 1] = rr[1];
      if (J) for (int k = 0; k < 2; ++k) {
        double* o = J + (size_t)i * 2 * nc + k * nc;
        if (nc == 13) for (int c = 0; c < 13; ++c) o[c] = JJ[13 * k + c];
        else if (nc == 6) for (int c = 0; c < 6; ++c) o[c] = JJ[13 * k + c];
        else o[0] = JJ[13 * k + 12];
      }
    }
  };
  if (n_threads <= 1) work(0, p->n_pobs);
  else { std::vector<std::thread> th; for (int t = 0; t < n_threads; ++t) th.emplace_back(work, (int)((long long)p->n_pobs * t / n_threads), (int)((long long)p->n_pobs * (t + 1) / n_threads)); for (auto& t : th) t.join(); }
  return 0;
}

extern "C" int tso_eval_text(int kind, int jac_mode, const tslam_ba_problem* p, double* r, double* J, int n_threads) {
  const int nc = kind == TSLAM_TX_BA ? 15 : (kind == TSLAM_TX_POSE ? 6 : 3);
  const unsigned mask = kind == TSLAM_TX_BA ? 7u : (kind == TSLAM_TX_POSE ? 1u : 4u);
  auto work = [&](int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      TextBlockConst c = text_const(*p, i);
      if (kind == TSLAM_TX_THETA) c.wT = 1.0;
      const double* cam = p->cams + 7 * (size_t)p->t_cam[i]; const double* host = p->cams + 7 * (size_t)p->t_host[i];
      const double* th = p->theta + 3 * (size_t)p->t_plane[i];
      double rr[8], JJ[120];
      if (!J) text_functor(cam, host, th, c, rr);
      else if (jac_mode == TSLAM_JAC_CENTRAL_DIFF) text_eval_numeric(cam, host, th, c, mask, rr, JJ);
      else text_eval_analytic(cam, host, th, c, rr, JJ);
      for (int k = 0; k < 8; ++k) r[8 * (size_t)i + k] = rr[k];
      if (J) for (int k = 0; k < 8; ++k) {
        double* o = J + (size_t)i * 8 * nc + k * nc;
        if (nc == 15) for (int q = 0; q < 15; ++q) o[q] = JJ[15 * k + q];
        else if (nc == 6) for (int q = 0; q < 6; ++q) o[q] = JJ[15 * k + q];
        else for (int q = 0; q < 3; ++q) o[q] = JJ[15 * k + 12 + q];
      }
    }
  };
  if (n_threads <= 1) work(0, p->n_tobs);
  else { std::vector<std::thread> th; for (int t = 0; t < n_threads; ++t) th.emplace_back(work, (int)((long long)p->n_tobs * t / n_threads), (int)((long long)p->n_tobs * (t + 1) / n_threads)); for (auto& t : th) t.join(); }
  return 0;
}

// Ambient (un-projected) 2x15 autodiff Jacobian of one point block — for Jacobian self-checks.
extern "C" int tso_point_ambient(const double* cam, const double* host, double rho, const double* ray_xy, const double* uv,
                                 const double* K4, const double* w, double* r, double* J30) {
  typedef Jet<15> JT;
  JT qc[4], tc[3], qh[4], th[3], jr(rho, 14), res[2];
  for (int i = 0; i < 4; ++i) { qc[i] = JT(cam[i], i); qh[i] = JT(host[i], 7 + i); }
  for (int i = 0; i < 3; ++i) { tc[i] = JT(cam[4 + i], 4 + i); th[i] = JT(host[4 + i], 11 + i); }
  const double ray[3] = {ray_xy[0], ray_xy[1], 1.0};
  point_functor<JT>(qc, tc, qh, th, jr, ray, uv, K4, w, res);
  for (int k = 0; k < 2; ++k) { r[k] = res[k].a; for (int c = 0; c < 15; ++c) J30[15 * k + c] = res[k].v[c]; }
  return 0;
}
extern "C" void tso_quat_plus(const double* x, const double* d, double* out) { quat_plus(x, d, out); }

// ceres::Covariance of the theta blocks (src/optimizer.cc:2219-2238): (J_theta' J_theta)^-1 per plane, loss-corrected Jacobian.
extern "C" int tso_theta_covariance(const tslam_ba_problem* p, int jac_mode, double* cov) {
  std::vector<double> V(9 * (size_t)p->n_planes, 0.0);
  for (int i = 0; i < p->n_tobs; ++i) {
    TextBlockConst c = text_const(*p, i);
    const double* cam = p->cams + 7 * (size_t)p->t_cam[i]; const double* host = p->cams + 7 * (size_t)p->t_host[i];
    const double* th = p->theta + 3 * (size_t)p->t_plane[i];
    double r[8], J[120];
    if (jac_mode == TSLAM_JAC_CENTRAL_DIFF) text_eval_numeric(cam, host, th, c, 7u, r, J); else text_eval_analytic(cam, host, th, c, r, J);
    double s = 0; for (int k = 0; k < 8; ++k) s += r[k] * r[k];
    double rho[3]; huber(p->huber_text, s, rho);
    double* Vp = &V[9 * (size_t)p->t_plane[i]];
    for (int k = 0; k < 8; ++k) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Vp[3 * a + b] += rho[1] * J[15 * k + 12 + a] * J[15 * k + 12 + b];
  }
  int singular = 0;
  for (int k = 0; k < p->n_planes; ++k) { if (!inv3_spd(&V[9 * (size_t)k], cov + 9 * (size_t)k)) { for (int a = 0; a < 9; ++a) cov[9 * (size_t)k + a] = 0; ++singular; } }
  return singular;
}
