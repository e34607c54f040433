// ORACLE — TEST INFRASTRUCTURE ONLY. Not linked, imported or executed by the product path
// (textslam_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// PARITY UNPINNED: the reference (SJTU-ViSYS/TextSLAM) ships no tests, golden vectors or fixtures
// for this path, and its arithmetic lives in Ceres Solver (version unpinned, CMakeLists.txt:7),
// Eigen and OpenCV, none of which are present in this image. This file restates
// the reference functors operation-for-operation in plain C++17 (double), with a forward-mode
// dual-number type standing in for ceres::Jet and a central-difference evaluator following
// Ceres' published NumericDiff step rule. A second restatement written from the same headers in plain numpy
// (tests/test_oracle_functors_py.py) agrees with it: residuals to 1e-11, Jacobians with central / complex-step derivatives
// through the quaternion Plus, the NumericDiff mode with Ceres' step rule applied to the ambient parameters.
//
// Reference anchors (all under /root/reference):
//   include/rotation.h:524-573         UnitQuaternionRotatePoint / QuaternionRotatePoint / QuaternionProduct
//   include/auto_BAScene.h:27-87       weighted reprojection, params (q_cw,t_cw,q_rw,t_rw,rho)
//   include/auto_BASceneNW.h:27-84     same, unweighted
//   include/auto_PoseOptimScene.h:28-88 host pose + landmark constant
//   include/auto_RhoScene.h:27-64      only rho free
//   include/nume_BAText.h:28-94        8 photometric residuals through plane theta
//   include/nume_PoseOptimText.h:28-79 theta + host pose constant
//   include/nume_thetaText.h:28-74     only theta free, unweighted
//   include/ModelTool.hpp:164-171      TextProj (3-arg)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <algorithm>

namespace tso {

// ---------------------------------------------------------------------------------------------
// Forward-mode dual number (stands in for ceres::Jet<double,N>).
// ---------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT implicit
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi; return h;
}
template <int N> inline Jet<N> sqrt(const Jet<N>& f) { Jet<N> h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
inline double sqrt(double x) { return std::sqrt(x); }

// ---------------------------------------------------------------------------------------------
// rotation.h:524-573
// ---------------------------------------------------------------------------------------------
template <typename T> inline void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  const T t2 = q[0] * q[1];
  const T t3 = q[0] * q[2];
  const T t4 = q[0] * q[3];
  const T t5 = -q[1] * q[1];
  const T t6 = q[1] * q[2];
  const T t7 = q[1] * q[3];
  const T t8 = -q[2] * q[2];
  const T t9 = q[2] * q[3];
  const T t1 = -q[3] * q[3];
  result[0] = T(2) * ((t8 + t1) * pt[0] + (t6 - t4) * pt[1] + (t3 + t7) * pt[2]) + pt[0];
  result[1] = T(2) * ((t4 + t6) * pt[0] + (t5 + t1) * pt[1] + (t9 - t2) * pt[2]) + pt[1];
  result[2] = T(2) * ((t7 - t3) * pt[0] + (t2 + t9) * pt[1] + (t5 + t8) * pt[2]) + pt[2];
}
template <typename T> inline void QuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  using std::sqrt;
  const T scale = T(1) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T unit[4] = {scale * q[0], scale * q[1], scale * q[2], scale * q[3]};
  UnitQuaternionRotatePoint(unit, pt, result);
}
template <typename T> inline void QuaternionProduct(const T z[4], const T w[4], T zw[4]) {
  zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}

// ---------------------------------------------------------------------------------------------
// Point reprojection functor (auto_BAScene.h:27-87). The four reference functors are the same
// geometry with different blocks held constant; `T` is double or Jet.
//   qcw,tcw : observing keyframe pose; qrw,trw : host keyframe pose; rho : inverse depth
//   ray = (x, y, 1) (src/mapPts.cc:54-61), K4 = fx,fy,cx,cy, w = residual weights.
// ---------------------------------------------------------------------------------------------
template <typename T>
inline void point_functor(const T qcw_[4], const T tcw_[3], const T qrw_[4], const T trw_[3], const T& rho,
                          const double ray[3], const double uv[2], const double K4[4], const double w[2], T res[2]) {
  T fx = T(K4[0]), fy = T(K4[1]), cx = T(K4[2]), cy = T(K4[3]);
  T qwr[4], qcw[4], qcr[4];
  qwr[0] = qrw_[0]; qwr[1] = -qrw_[1]; qwr[2] = -qrw_[2]; qwr[3] = -qrw_[3];
  qcw[0] = qcw_[0]; qcw[1] = qcw_[1]; qcw[2] = qcw_[2]; qcw[3] = qcw_[3];
  qcr[0] = qcw[0] * qwr[0] - qcw[1] * qwr[1] - qcw[2] * qwr[2] - qcw[3] * qwr[3];
  qcr[1] = qcw[0] * qwr[1] + qcw[1] * qwr[0] + qcw[2] * qwr[3] - qcw[3] * qwr[2];
  qcr[2] = qcw[0] * qwr[2] - qcw[1] * qwr[3] + qcw[2] * qwr[0] + qcw[3] * qwr[1];
  qcr[3] = qcw[0] * qwr[3] + qcw[1] * qwr[2] - qcw[2] * qwr[1] + qcw[3] * qwr[0];
  T trw[3] = {trw_[0], trw_[1], trw_[2]}, tcr[3], tcr_tmp[3];
  QuaternionRotatePoint(qcr, trw, tcr_tmp);
  tcr[0] = -tcr_tmp[0] + tcw_[0];
  tcr[1] = -tcr_tmp[1] + tcw_[1];
  tcr[2] = -tcr_tmp[2] + tcw_[2];
  T p[3];
  p[0] = T(1.0) / rho * T(ray[0]);
  p[1] = T(1.0) / rho * T(ray[1]);
  p[2] = T(1.0) / rho * T(ray[2]);
  T qp[3];
  QuaternionRotatePoint(qcr, p, qp);
  T u = fx * (qp[0] + tcr[0]) / (qp[2] + tcr[2]) + cx;
  T v = fy * (qp[1] + tcr[1]) / (qp[2] + tcr[2]) + cy;
  res[0] = (u - T(uv[0])) * T(w[0]);
  res[1] = (v - T(uv[1])) * T(w[1]);
}

// Ceres QuaternionParameterization::ComputeJacobian (4x3, row-major) — SURVEY Appendix A.1.
inline void quat_local_jacobian(const double x[4], double P[12]) {
  P[0] = -x[1]; P[1] = -x[2]; P[2] = -x[3];
  P[3] = x[0];  P[4] = x[3];  P[5] = -x[2];
  P[6] = -x[3]; P[7] = x[0];  P[8] = x[1];
  P[9] = x[2];  P[10] = -x[1]; P[11] = x[0];
}
// Ceres QuaternionParameterization::Plus: x_plus = dq(delta) (x) x, no renormalisation.
inline void quat_plus(const double x[4], const double d[3], double out[4]) {
  const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double s = std::sin(n) / n;
    double dq[4] = {std::cos(n), s * d[0], s * d[1], s * d[2]};
    QuaternionProduct(dq, x, out);
  } else {
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
  }
}

// Tangent-space (Ceres "local") residual + Jacobian of one point block.
//   free_mask bit0: observing cam, bit1: host cam, bit2: rho. J is 2x13 row-major with columns
//   [d_c(3) t_c(3) d_h(3) t_h(3) rho]; columns of constant blocks are left 0.
inline void point_eval(const double* cam, const double* host, double rho, const double ray[3], const double uv[2],
                       const double K4[4], const double w[2], double r[2], double* J /*26 or null*/) {
  if (!J) {
    point_functor<double>(cam, cam + 4, host, host + 4, rho, ray, uv, K4, w, r);
    return;
  }
  typedef Jet<15> JT;
  JT qc[4], tc[3], qh[4], th[3], jr(rho, 14), res[2];
  for (int i = 0; i < 4; ++i) { qc[i] = JT(cam[i], i); qh[i] = JT(host[i], 7 + i); }
  for (int i = 0; i < 3; ++i) { tc[i] = JT(cam[4 + i], 4 + i); th[i] = JT(host[4 + i], 11 + i); }
  point_functor<JT>(qc, tc, qh, th, jr, ray, uv, K4, w, res);
  double Pc[12], Ph[12];
  quat_local_jacobian(cam, Pc);
  quat_local_jacobian(host, Ph);
  for (int k = 0; k < 2; ++k) {
    r[k] = res[k].a;
    double* Jr = J + 13 * k;
    const double* a = res[k].v;
    for (int c = 0; c < 3; ++c) {
      Jr[c] = a[0] * Pc[c] + a[1] * Pc[3 + c] + a[2] * Pc[6 + c] + a[3] * Pc[9 + c];
      Jr[3 + c] = a[4 + c];
      Jr[6 + c] = a[7] * Ph[c] + a[8] * Ph[3 + c] + a[9] * Ph[6 + c] + a[10] * Ph[9 + c];
      Jr[9 + c] = a[11 + c];
    }
    Jr[12] = a[14];
  }
}

// ---------------------------------------------------------------------------------------------
// Text functor (nume_BAText.h:28-94): quaternion -> rotation matrix exactly as
// Eigen::Quaterniond::normalized().toRotationMatrix() computes it, Tcr = Tcw * Trw^-1 with the
// rigid-transform inverse written out (Eigen's general 4x4 inverse differs only in rounding).
// ---------------------------------------------------------------------------------------------
inline void quat_to_R_eigen(const double q_[4], double R[9]) {
  const double n = std::sqrt(q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]);
  const double w = q_[0] / n, x = q_[1] / n, y = q_[2] / n, z = q_[3] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}
// Tcr = Tcw * Trw^{-1}:  Rcr = Rcw Rrw^T, tcr = tcw - Rcr trw.
inline void relative_pose(const double* cam, const double* host, double Rcr[9], double tcr[3]) {
  double Rc[9], Rh[9];
  quat_to_R_eigen(cam, Rc);
  quat_to_R_eigen(host, Rh);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Rcr[3 * i + j] = Rc[3 * i + 0] * Rh[3 * j + 0] + Rc[3 * i + 1] * Rh[3 * j + 1] + Rc[3 * i + 2] * Rh[3 * j + 2];
  for (int i = 0; i < 3; ++i)
    tcr[i] = cam[4 + i] - (Rcr[3 * i + 0] * host[4] + Rcr[3 * i + 1] * host[5] + Rcr[3 * i + 2] * host[6]);
}

struct TextBlockConst {
  const uint8_t* img; int cols, rows;   // stride == cols (nume_BAText.h:25)
  const double* rays;                   // 8 x (x,y), z == 1 (src/tool.cc:1550-1561)
  const double* iref;                   // 8 normalised reference intensities
  double mu, sigma;
  const double* K4;                     // level intrinsics
  double wT;
};

// One functor call: 8 residuals. (nume_BAText.h:54-91; TextProj ModelTool.hpp:164-171)
inline void text_functor(const double* cam, const double* host, const double theta[3], const TextBlockConst& c, double res[8]) {
  double R[9], t[3];
  relative_pose(cam, host, R, t);
  for (int i = 0; i < 8; ++i) {
    const double rx = c.rays[2 * i], ry = c.rays[2 * i + 1], rz = 1.0;
    const double rho = -(rx * theta[0] + ry * theta[1] + rz * theta[2]);
    const double px = (R[0] * rx + R[1] * ry + R[2] * rz) / rho + t[0];
    const double py = (R[3] * rx + R[4] * ry + R[5] * rz) / rho + t[1];
    const double pz = (R[6] * rx + R[7] * ry + R[8] * rz) / rho + t[2];
    const double u = c.K4[0] * px / pz + c.K4[2];
    const double v = c.K4[1] * py / pz + c.K4[3];
    double inten;
    const int uf = (int)std::floor(u), vf = (int)std::floor(v), uc = (int)std::ceil(u), vc = (int)std::ceil(v);
    if (!(u == u) || !(v == v) || uf < 0 || vf < 0 || uc >= c.cols || vc >= c.rows) {
      inten = 0;
    } else {
      const uint8_t* p = c.img + (size_t)vf * c.cols + uf;
      const double su = u - uf, sv = v - vf;
      const double wtl = (1.0 - su) * (1.0 - sv), wtr = su * (1.0 - sv), wbl = (1.0 - su) * sv, wbr = su * sv;
      // when u (v) is integral uc==uf and the +1 tap has weight 0; guard the read at the last column/row
      const int du = (uf + 1 < c.cols) ? 1 : 0, dv = (vf + 1 < c.rows) ? c.cols : 0;
      inten = wtl * p[0] + wtr * p[du] + wbl * p[dv] + wbr * p[dv + du];
    }
    if (c.sigma != 0) {
      const double n = (inten - c.mu) / c.sigma;
      res[i] = (n - c.iref[i]) * c.wT;
    } else {
      res[i] = 0.0;
    }
  }
}

// Ceres NumericDiff CENTRAL, default options:
//   h_j = max(|x_j| * 1e-6, sqrt(DBL_EPSILON)); col_j = (f(x+h) - f(x-h)) * ((1/h)/2)
// Ambient layout x = [qc(4) tc(3) qh(4) th(3) theta(3)] (17). J_amb is 8x17 row-major.
inline void text_numeric_ambient(const double* cam, const double* host, const double theta[3], const TextBlockConst& c,
                                 unsigned free_mask, double r[8], double Jamb[8 * 17]) {
  double x[17];
  std::memcpy(x, cam, 7 * sizeof(double));
  std::memcpy(x + 7, host, 7 * sizeof(double));
  std::memcpy(x + 14, theta, 3 * sizeof(double));
  text_functor(x, x + 7, x + 14, c, r);
  std::memset(Jamb, 0, sizeof(double) * 8 * 17);
  const double min_step = std::sqrt(DBL_EPSILON);
  for (int j = 0; j < 17; ++j) {
    const int blk = j < 7 ? 0 : (j < 14 ? 1 : 2);
    if (!(free_mask & (1u << blk))) continue;
    const double xj = x[j];
    const double delta = std::max(min_step, std::fabs(xj) * 1e-6);
    double fp[8], fm[8];
    x[j] = xj + delta; text_functor(x, x + 7, x + 14, c, fp);
    x[j] = xj - delta; text_functor(x, x + 7, x + 14, c, fm);
    x[j] = xj;
    const double one_over = (1.0 / delta) / 2;
    for (int i = 0; i < 8; ++i) Jamb[17 * i + j] = (fp[i] - fm[i]) * one_over;
  }
}

// Tangent-space 8x15 Jacobian [d_c t_c d_h t_h theta] from the ambient numeric one.
inline void text_eval_numeric(const double* cam, const double* host, const double theta[3], const TextBlockConst& c,
                              unsigned free_mask, double r[8], double* J /*8x15 or null*/) {
  if (!J) { text_functor(cam, host, theta, c, r); return; }
  double Ja[8 * 17];
  text_numeric_ambient(cam, host, theta, c, free_mask, r, Ja);
  double Pc[12], Ph[12];
  quat_local_jacobian(cam, Pc);
  quat_local_jacobian(host, Ph);
  for (int i = 0; i < 8; ++i) {
    const double* a = Ja + 17 * i;
    double* Jr = J + 15 * i;
    for (int k = 0; k < 3; ++k) {
      Jr[k] = a[0] * Pc[k] + a[1] * Pc[3 + k] + a[2] * Pc[6 + k] + a[3] * Pc[9 + k];
      Jr[3 + k] = a[4 + k];
      Jr[6 + k] = a[7] * Ph[k] + a[8] * Ph[3 + k] + a[9] * Ph[6 + k] + a[10] * Ph[9 + k];
      Jr[9 + k] = a[11 + k];
      Jr[12 + k] = a[14 + k];
    }
  }
}

// Analytic tangent-space Jacobian — used to cross-validate the numeric one
// and as the oracle for the GPU "analytic" jac_mode.
inline void text_eval_analytic(const double* cam, const double* host, const double theta[3], const TextBlockConst& c,
                               double r[8], double* J /*8x15*/) {
  double R[9], t[3];
  relative_pose(cam, host, R, t);
  for (int i = 0; i < 8; ++i) {
    const double ray[3] = {c.rays[2 * i], c.rays[2 * i + 1], 1.0};
    const double rho = -(ray[0] * theta[0] + ray[1] * theta[1] + ray[2] * theta[2]);
    double Rr[3], X[3], pc[3];
    for (int k = 0; k < 3; ++k) Rr[k] = R[3 * k] * ray[0] + R[3 * k + 1] * ray[1] + R[3 * k + 2] * ray[2];
    for (int k = 0; k < 3; ++k) { pc[k] = Rr[k] / rho + t[k]; X[k] = pc[k] - cam[4 + k]; }
    const double u = c.K4[0] * pc[0] / pc[2] + c.K4[2];
    const double v = c.K4[1] * pc[1] / pc[2] + c.K4[3];
    double inten = 0, gu = 0, gv = 0;
    const int uf = (int)std::floor(u), vf = (int)std::floor(v), uc = (int)std::ceil(u), vc = (int)std::ceil(v);
    if (!(!(u == u) || !(v == v) || uf < 0 || vf < 0 || uc >= c.cols || vc >= c.rows)) {
      const uint8_t* p = c.img + (size_t)vf * c.cols + uf;
      const double su = u - uf, sv = v - vf;
      const int du = (uf + 1 < c.cols) ? 1 : 0, dv = (vf + 1 < c.rows) ? c.cols : 0;
      const double I00 = p[0], I01 = p[du], I10 = p[dv], I11 = p[dv + du];
      inten = (1.0 - su) * (1.0 - sv) * I00 + su * (1.0 - sv) * I01 + (1.0 - su) * sv * I10 + su * sv * I11;
      gu = (1.0 - sv) * (I01 - I00) + sv * (I11 - I10);
      gv = (1.0 - su) * (I10 - I00) + su * (I11 - I01);
    }
    double* Jr = J + 15 * i;
    if (c.sigma != 0) {
      r[i] = ((inten - c.mu) / c.sigma - c.iref[i]) * c.wT;
      const double s = c.wT / c.sigma;
      const double iz = 1.0 / pc[2];
      // d r / d pc
      const double a0 = s * gu * c.K4[0] * iz;
      const double a1 = s * gv * c.K4[1] * iz;
      const double a2 = -s * (gu * c.K4[0] * pc[0] + gv * c.K4[1] * pc[1]) * iz * iz;
      const double a[3] = {a0, a1, a2};
      // d pc / d delta_c = -2 [X]x  =>  row a * (-2[X]x) = -2 (a x X)^T ... written out:
      // [X]x = [[0,-X2,X1],[X2,0,-X0],[-X1,X0,0]];  a^T [X]x = (a1 X2 - a2 X1, a2 X0 - a0 X2, a0 X1 - a1 X0)
      Jr[0] = -2.0 * (a[1] * X[2] - a[2] * X[1]);
      Jr[1] = -2.0 * (a[2] * X[0] - a[0] * X[2]);
      Jr[2] = -2.0 * (a[0] * X[1] - a[1] * X[0]);
      Jr[3] = a[0]; Jr[4] = a[1]; Jr[5] = a[2];
      // b = a^T R_cr ; d pc/d delta_h = 2 R_cr [p_r - t_rw]x ; d pc / d t_rw = -R_cr
      double b[3], m[3];
      for (int k = 0; k < 3; ++k) b[k] = a[0] * R[k] + a[1] * R[3 + k] + a[2] * R[6 + k];
      for (int k = 0; k < 3; ++k) m[k] = ray[k] / rho - host[4 + k];
      Jr[6] = 2.0 * (b[1] * m[2] - b[2] * m[1]);
      Jr[7] = 2.0 * (b[2] * m[0] - b[0] * m[2]);
      Jr[8] = 2.0 * (b[0] * m[1] - b[1] * m[0]);
      Jr[9] = -b[0]; Jr[10] = -b[1]; Jr[11] = -b[2];
      // d pc / d theta = R ray ray^T / rho^2
      const double aRr = (a[0] * Rr[0] + a[1] * Rr[1] + a[2] * Rr[2]) / (rho * rho);
      Jr[12] = aRr * ray[0]; Jr[13] = aRr * ray[1]; Jr[14] = aRr * ray[2];
    } else {
      r[i] = 0.0;
      for (int k = 0; k < 15; ++k) Jr[k] = 0.0;
    }
  }
}

// Huber loss (ceres::HuberLoss(a)) on s = ||r||^2; returns rho[0..2]. a <= 0 => trivial loss.
inline void huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (a <= 0.0 || s <= b) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; return; }
  const double r = std::sqrt(s);
  rho[0] = 2.0 * a * r - b;
  rho[1] = std::max(DBL_MIN, a / r);
  rho[2] = -rho[1] / (2.0 * s);
}

}  // namespace tso
