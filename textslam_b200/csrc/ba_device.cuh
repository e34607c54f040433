// Device-side math of the BA hot path (sm_100a). FP64 throughout (SURVEY §7 hard part 4).
//
// Residual formulas follow the reference functors' operation order so that values agree with the
// CPU oracle to rounding: include/auto_BAScene.h:43-84 (points), include/nume_BAText.h:35-91 +
// include/ModelTool.hpp:164-171 (text). Jacobians are closed-form tangent-space derivatives
// consistent with ceres::QuaternionParameterization::Plus instead of
// Jets: d p_c/d delta_c = -2[X]x, d/d t_c = I, d/d delta_h = 2 R_cr [p_r - t_h]x, d/d t_h = -R_cr,
// d/d rho = -R_cr ray / rho^2, d/d theta = R_cr ray ray^T / rho_theta^2.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>

namespace tsl {

struct Cam {  // T_cw
  double q[4];
  double t[3];
};

__device__ __forceinline__ Cam load_cam(const double* __restrict__ cams, int k) {
  Cam c;
  const double* p = cams + 7 * (size_t)k;
#pragma unroll
  for (int i = 0; i < 4; ++i) c.q[i] = __ldg(p + i);
#pragma unroll
  for (int i = 0; i < 3; ++i) c.t[i] = __ldg(p + 4 + i);
  return c;
}

// include/rotation.h:525-541 with the normalisation of :543-562 folded in by the caller.
__device__ __forceinline__ void unit_quat_rotate(const double q[4], const double pt[3], double r[3]) {
  const double t2 = q[0] * q[1], t3 = q[0] * q[2], t4 = q[0] * q[3];
  const double t5 = -q[1] * q[1], t6 = q[1] * q[2], t7 = q[1] * q[3];
  const double t8 = -q[2] * q[2], t9 = q[2] * q[3], t1 = -q[3] * q[3];
  r[0] = 2.0 * ((t8 + t1) * pt[0] + (t6 - t4) * pt[1] + (t3 + t7) * pt[2]) + pt[0];
  r[1] = 2.0 * ((t4 + t6) * pt[0] + (t5 + t1) * pt[1] + (t9 - t2) * pt[2]) + pt[1];
  r[2] = 2.0 * ((t7 - t3) * pt[0] + (t2 + t9) * pt[1] + (t5 + t8) * pt[2]) + pt[2];
}

// rotation matrix (row-major) of a UNIT quaternion, same entries unit_quat_rotate applies.
__device__ __forceinline__ void unit_quat_to_R(const double q[4], double R[9]) {
  const double t2 = q[0] * q[1], t3 = q[0] * q[2], t4 = q[0] * q[3];
  const double t5 = -q[1] * q[1], t6 = q[1] * q[2], t7 = q[1] * q[3];
  const double t8 = -q[2] * q[2], t9 = q[2] * q[3], t1 = -q[3] * q[3];
  R[0] = 2.0 * (t8 + t1) + 1.0; R[1] = 2.0 * (t6 - t4);       R[2] = 2.0 * (t3 + t7);
  R[3] = 2.0 * (t4 + t6);       R[4] = 2.0 * (t5 + t1) + 1.0; R[5] = 2.0 * (t9 - t2);
  R[6] = 2.0 * (t7 - t3);       R[7] = 2.0 * (t2 + t9);       R[8] = 2.0 * (t5 + t8) + 1.0;
}

// q_cr = q_cw (x) conj(q_rw), normalised (auto_BAScene.h:43-55 + rotation.h:547-558).
__device__ __forceinline__ void relative_unit_quat(const double qc[4], const double qh[4], double u[4]) {
  const double w0 = qh[0], w1 = -qh[1], w2 = -qh[2], w3 = -qh[3];
  double q0 = qc[0] * w0 - qc[1] * w1 - qc[2] * w2 - qc[3] * w3;
  double q1 = qc[0] * w1 + qc[1] * w0 + qc[2] * w3 - qc[3] * w2;
  double q2 = qc[0] * w2 - qc[1] * w3 + qc[2] * w0 + qc[3] * w1;
  double q3 = qc[0] * w3 + qc[1] * w2 - qc[2] * w1 + qc[3] * w0;
  const double scale = 1.0 / sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  u[0] = scale * q0; u[1] = scale * q1; u[2] = scale * q2; u[3] = scale * q3;
}

// ------------------------------------------------------------------------------------------------
// Point block: residual (2) + tangent Jacobian 2x13 [d_c t_c d_h t_h rho].
// ------------------------------------------------------------------------------------------------
template <bool WANT_J>
__device__ __forceinline__ void point_eval(const Cam& c, const Cam& h, double rho, double rx, double ry, double u_obs,
                                           double v_obs, double fx, double fy, double cx, double cy, double wx, double wy,
                                           double r[2], double J[26]) {
  double qcr[4];
  relative_unit_quat(c.q, h.q, qcr);
  double tmp[3], qp[3];
  unit_quat_rotate(qcr, h.t, tmp);                       // R_cr t_rw
  const double tcr0 = -tmp[0] + c.t[0], tcr1 = -tmp[1] + c.t[1], tcr2 = -tmp[2] + c.t[2];
  const double ir = 1.0 / rho;
  const double p[3] = {ir * rx, ir * ry, ir * 1.0};
  unit_quat_rotate(qcr, p, qp);                          // R_cr p_r
  const double x = qp[0] + tcr0, y = qp[1] + tcr1, z = qp[2] + tcr2;
  const double u = fx * x / z + cx;
  const double v = fy * y / z + cy;
  r[0] = (u - u_obs) * wx;
  r[1] = (v - v_obs) * wy;
  if (WANT_J) {
    const double iz = 1.0 / z;
    // rows of d(r)/d(p_c)
    const double a0[3] = {wx * fx * iz, 0.0, -wx * fx * x * iz * iz};
    const double a1[3] = {0.0, wy * fy * iz, -wy * fy * y * iz * iz};
    const double X[3] = {qp[0] - tmp[0], qp[1] - tmp[1], qp[2] - tmp[2]};  // R_cr (p_r - t_rw)
    double R[9];
    unit_quat_to_R(qcr, R);
    const double m[3] = {p[0] - h.t[0], p[1] - h.t[1], p[2] - h.t[2]};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double* a = k == 0 ? a0 : a1;
      double* Jr = J + 13 * k;
      // a^T (-2 [X]x) = -2 (a x X)
      Jr[0] = -2.0 * (a[1] * X[2] - a[2] * X[1]);
      Jr[1] = -2.0 * (a[2] * X[0] - a[0] * X[2]);
      Jr[2] = -2.0 * (a[0] * X[1] - a[1] * X[0]);
      Jr[3] = a[0]; Jr[4] = a[1]; Jr[5] = a[2];
      const double b0 = a[0] * R[0] + a[1] * R[3] + a[2] * R[6];
      const double b1 = a[0] * R[1] + a[1] * R[4] + a[2] * R[7];
      const double b2 = a[0] * R[2] + a[1] * R[5] + a[2] * R[8];
      Jr[6] = 2.0 * (b1 * m[2] - b2 * m[1]);
      Jr[7] = 2.0 * (b2 * m[0] - b0 * m[2]);
      Jr[8] = 2.0 * (b0 * m[1] - b1 * m[0]);
      Jr[9] = -b0; Jr[10] = -b1; Jr[11] = -b2;
      Jr[12] = -(a[0] * qp[0] + a[1] * qp[1] + a[2] * qp[2]) * ir;  // -a . R_cr ray / rho^2
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Text: relative pose the way nume_BAText.h:35-51 builds it (Eigen normalized().toRotationMatrix()).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quat_to_R_eigen(const double q_[4], double R[9]) {
  const double n = sqrt(q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]);
  const double w = q_[0] / n, x = q_[1] / n, y = q_[2] / n, z = q_[3] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}
struct RelPose { double R[9]; double t[3]; };
__device__ __forceinline__ void relative_pose(const double qc[4], const double tc[3], const double qh[4], const double th[3], RelPose& P) {
  double Rc[9], Rh[9];
  quat_to_R_eigen(qc, Rc);
  quat_to_R_eigen(qh, Rh);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      P.R[3 * i + j] = Rc[3 * i + 0] * Rh[3 * j + 0] + Rc[3 * i + 1] * Rh[3 * j + 1] + Rc[3 * i + 2] * Rh[3 * j + 2];
#pragma unroll
  for (int i = 0; i < 3; ++i) P.t[i] = tc[i] - (P.R[3 * i + 0] * th[0] + P.R[3 * i + 1] * th[1] + P.R[3 * i + 2] * th[2]);
}

// tile != nullptr: the taps come from a shared-memory copy of the image window [tx0, tx0 + tw) x [ty0, ..) that the CTA staged
// with a TMA tensor load (ba_eval_tma.cu); otherwise from global memory through the read-only path.
struct TextImg { const uint8_t* img; int cols, rows; const uint8_t* tile = nullptr; int tx0 = 0, ty0 = 0, tw = 0; };

// One pixel of the 8-pixel pattern: intensity + bilinear gradient (nume_BAText.h:58-91).
template <bool WANT_G>
__device__ __forceinline__ double text_pixel(const RelPose& P, double rx, double ry, const double th[3], const TextImg& im,
                                             double fx, double fy, double cx, double cy, double pc[3], double Rr[3],
                                             double& rho, double& gu, double& gv) {
  rho = -(rx * th[0] + ry * th[1] + 1.0 * th[2]);
  Rr[0] = P.R[0] * rx + P.R[1] * ry + P.R[2] * 1.0;
  Rr[1] = P.R[3] * rx + P.R[4] * ry + P.R[5] * 1.0;
  Rr[2] = P.R[6] * rx + P.R[7] * ry + P.R[8] * 1.0;
  pc[0] = Rr[0] / rho + P.t[0];
  pc[1] = Rr[1] / rho + P.t[1];
  pc[2] = Rr[2] / rho + P.t[2];
  const double u = fx * pc[0] / pc[2] + cx;
  const double v = fy * pc[1] / pc[2] + cy;
  const double ufl = floor(u), vfl = floor(v);
  double inten = 0.0;
  gu = 0.0; gv = 0.0;
  // uf<0 || vf<0 || uc>=cols || vc>=rows -> 0 (NaN falls through to 0 as well)
  if (ufl >= 0.0 && vfl >= 0.0 && ceil(u) < (double)im.cols && ceil(v) < (double)im.rows) {
    const int uf = (int)ufl, vf = (int)vfl;
    const double su = u - ufl, sv = v - vfl;
    double I00, I01, I10, I11;
    if (im.tile) {
      const uint8_t* p = im.tile + (vf - im.ty0) * im.tw + (uf - im.tx0);
      const int du = (uf + 1 < im.cols) ? 1 : 0, dv = (vf + 1 < im.rows) ? im.tw : 0;
      I00 = (double)p[0]; I01 = (double)p[du]; I10 = (double)p[dv]; I11 = (double)p[dv + du];
    } else {
      const uint8_t* p = im.img + (size_t)vf * im.cols + uf;
      const int du = (uf + 1 < im.cols) ? 1 : 0, dv = (vf + 1 < im.rows) ? im.cols : 0;
      I00 = (double)__ldg(p); I01 = (double)__ldg(p + du); I10 = (double)__ldg(p + dv); I11 = (double)__ldg(p + dv + du);
    }
    const double wtl = (1.0 - su) * (1.0 - sv), wtr = su * (1.0 - sv), wbl = (1.0 - su) * sv, wbr = su * sv;
    inten = wtl * I00 + wtr * I01 + wbl * I10 + wbr * I11;
    if (WANT_G) {
      gu = (1.0 - sv) * (I01 - I00) + sv * (I11 - I10);
      gv = (1.0 - su) * (I10 - I00) + su * (I11 - I01);
    }
  }
  return inten;
}

__device__ __forceinline__ double text_residual_only(const double qc[4], const double tc[3], const double qh[4], const double th_[3],
                                                     const double theta[3], double rx, double ry, const TextImg& im,
                                                     double fx, double fy, double cx, double cy, double mu, double sigma,
                                                     double iref, double wT) {
  if (sigma == 0.0) return 0.0;
  RelPose P;
  relative_pose(qc, tc, qh, th_, P);
  double pc[3], Rr[3], rho, gu, gv;
  const double inten = text_pixel<false>(P, rx, ry, theta, im, fx, fy, cx, cy, pc, Rr, rho, gu, gv);
  return ((inten - mu) / sigma - iref) * wT;
}

// residual + analytic tangent Jacobian row (15) of one pattern pixel
__device__ __forceinline__ double text_pixel_analytic(const Cam& c, const Cam& h, const double theta[3], double rx, double ry,
                                                      const TextImg& im, double fx, double fy, double cx, double cy, double mu,
                                                      double sigma, double iref, double wT, double Jr[15]) {
  if (sigma == 0.0) {
#pragma unroll
    for (int k = 0; k < 15; ++k) Jr[k] = 0.0;
    return 0.0;
  }
  RelPose P;
  relative_pose(c.q, c.t, h.q, h.t, P);
  double pc[3], Rr[3], rho, gu, gv;
  const double inten = text_pixel<true>(P, rx, ry, theta, im, fx, fy, cx, cy, pc, Rr, rho, gu, gv);
  const double res = ((inten - mu) / sigma - iref) * wT;
  const double s = wT / sigma;
  const double iz = 1.0 / pc[2];
  const double a[3] = {s * gu * fx * iz, s * gv * fy * iz, -s * (gu * fx * pc[0] + gv * fy * pc[1]) * iz * iz};
  const double X[3] = {pc[0] - c.t[0], pc[1] - c.t[1], pc[2] - c.t[2]};
  Jr[0] = -2.0 * (a[1] * X[2] - a[2] * X[1]);
  Jr[1] = -2.0 * (a[2] * X[0] - a[0] * X[2]);
  Jr[2] = -2.0 * (a[0] * X[1] - a[1] * X[0]);
  Jr[3] = a[0]; Jr[4] = a[1]; Jr[5] = a[2];
  const double b0 = a[0] * P.R[0] + a[1] * P.R[3] + a[2] * P.R[6];
  const double b1 = a[0] * P.R[1] + a[1] * P.R[4] + a[2] * P.R[7];
  const double b2 = a[0] * P.R[2] + a[1] * P.R[5] + a[2] * P.R[8];
  const double m[3] = {rx / rho - h.t[0], ry / rho - h.t[1], 1.0 / rho - h.t[2]};
  Jr[6] = 2.0 * (b1 * m[2] - b2 * m[1]);
  Jr[7] = 2.0 * (b2 * m[0] - b0 * m[2]);
  Jr[8] = 2.0 * (b0 * m[1] - b1 * m[0]);
  Jr[9] = -b0; Jr[10] = -b1; Jr[11] = -b2;
  const double aRr = (a[0] * Rr[0] + a[1] * Rr[1] + a[2] * Rr[2]) / (rho * rho);
  Jr[12] = aRr * rx; Jr[13] = aRr * ry; Jr[14] = aRr * 1.0;
  return res;
}

// The functor exactly as Ceres' NumericDiff evaluates it: one residual of nume_BAText::operator() (include/nume_BAText.h:28-94)
// with the reference's own operation sequence — Eigen's normalized().toRotationMatrix(), Tcr = Tcw Trw^-1, TextProj
// (include/ModelTool.hpp:164-171), bilinear taps — every operation an explicitly rounded IEEE double operation (no FMA
// contraction). Central differences divide the difference of two such values by 2h ~ 3e-8: a last-bit difference between two
// formulations of the same functor becomes a 1e-6..1e-5 relative difference of a Jacobian entry, so the Ceres-faithful mode
// evaluates the SAME arithmetic as the CPU restatement (oracle/ba_math.hpp text_functor, built with -ffp-contract=off).
__device__ __forceinline__ void quat_to_R_exact(const double q[4], double R[9]) {
  const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], q[0]), __dmul_rn(q[1], q[1])), __dmul_rn(q[2], q[2])), __dmul_rn(q[3], q[3])));
  const double w = __ddiv_rn(q[0], n), x = __ddiv_rn(q[1], n), y = __ddiv_rn(q[2], n), z = __ddiv_rn(q[3], n);
  const double tx = __dmul_rn(2.0, x), ty = __dmul_rn(2.0, y), tz = __dmul_rn(2.0, z);
  const double twx = __dmul_rn(tx, w), twy = __dmul_rn(ty, w), twz = __dmul_rn(tz, w);
  const double txx = __dmul_rn(tx, x), txy = __dmul_rn(ty, x), txz = __dmul_rn(tz, x);
  const double tyy = __dmul_rn(ty, y), tyz = __dmul_rn(tz, y), tzz = __dmul_rn(tz, z);
  R[0] = __dsub_rn(1.0, __dadd_rn(tyy, tzz)); R[1] = __dsub_rn(txy, twz);                 R[2] = __dadd_rn(txz, twy);
  R[3] = __dadd_rn(txy, twz);                 R[4] = __dsub_rn(1.0, __dadd_rn(txx, tzz)); R[5] = __dsub_rn(tyz, twx);
  R[6] = __dsub_rn(txz, twy);                 R[7] = __dadd_rn(tyz, twx);                 R[8] = __dsub_rn(1.0, __dadd_rn(txx, tyy));
}
__device__ __forceinline__ double dot3_exact(double a0, double b0, double a1, double b1, double a2, double b2) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}
static __device__ __noinline__ double text_residual_exact(const double* x /*[qc tc qh th theta] (17)*/, double rx, double ry, const TextImg& im,
                                                   double fx, double fy, double cx, double cy, double mu, double sigma, double iref, double wT) {
  if (sigma == 0.0) return 0.0;
  double Rc[9], Rh[9], R[9], t[3];
  quat_to_R_exact(x, Rc);
  quat_to_R_exact(x + 7, Rh);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[3 * i + j] = dot3_exact(Rc[3 * i], Rh[3 * j], Rc[3 * i + 1], Rh[3 * j + 1], Rc[3 * i + 2], Rh[3 * j + 2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = __dsub_rn(x[4 + i], dot3_exact(R[3 * i], x[11], R[3 * i + 1], x[12], R[3 * i + 2], x[13]));
  const double* th = x + 14;
  const double rho = -dot3_exact(rx, th[0], ry, th[1], 1.0, th[2]);
  const double px = __dadd_rn(__ddiv_rn(dot3_exact(R[0], rx, R[1], ry, R[2], 1.0), rho), t[0]);
  const double py = __dadd_rn(__ddiv_rn(dot3_exact(R[3], rx, R[4], ry, R[5], 1.0), rho), t[1]);
  const double pz = __dadd_rn(__ddiv_rn(dot3_exact(R[6], rx, R[7], ry, R[8], 1.0), rho), t[2]);
  const double u = __dadd_rn(__ddiv_rn(__dmul_rn(fx, px), pz), cx);
  const double v = __dadd_rn(__ddiv_rn(__dmul_rn(fy, py), pz), cy);
  const double ufl = floor(u), vfl = floor(v);
  double inten = 0.0;
  if (ufl >= 0.0 && vfl >= 0.0 && ceil(u) < (double)im.cols && ceil(v) < (double)im.rows) {
    const int uf = (int)ufl, vf = (int)vfl;
    const uint8_t* p = im.img + (size_t)vf * im.cols + uf;
    const double su = __dsub_rn(u, ufl), sv = __dsub_rn(v, vfl);
    const int du = (uf + 1 < im.cols) ? 1 : 0, dv = (vf + 1 < im.rows) ? im.cols : 0;
    const double I00 = (double)__ldg(p), I01 = (double)__ldg(p + du), I10 = (double)__ldg(p + dv), I11 = (double)__ldg(p + dv + du);
    const double osu = __dsub_rn(1.0, su), osv = __dsub_rn(1.0, sv);
    const double wtl = __dmul_rn(osu, osv), wtr = __dmul_rn(su, osv), wbl = __dmul_rn(osu, sv), wbr = __dmul_rn(su, sv);
    inten = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(wtl, I00), __dmul_rn(wtr, I01)), __dmul_rn(wbl, I10)), __dmul_rn(wbr, I11));
  }
  return __dmul_rn(__dsub_rn(__ddiv_rn(__dsub_rn(inten, mu), sigma), iref), wT);
}

// residual + Ceres NumericDiff<CENTRAL> replica projected to the tangent space.
// free_mask bit0 cam, bit1 host, bit2 theta (constant blocks get no Jacobian, like Ceres).
__device__ __forceinline__ double text_pixel_central(const Cam& c, const Cam& h, const double theta[3], double rx, double ry,
                                                     const TextImg& im, double fx, double fy, double cx, double cy, double mu,
                                                     double sigma, double iref, double wT, unsigned free_mask, double Jr[15]) {
  double x[17];
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[i] = c.q[i]; x[7 + i] = h.q[i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { x[4 + i] = c.t[i]; x[11 + i] = h.t[i]; x[14 + i] = theta[i]; }
  const double res = text_residual_exact(x, rx, ry, im, fx, fy, cx, cy, mu, sigma, iref, wT);
  double Ja[17];
  const double min_step = 1.4901161193847656e-08;  // sqrt(DBL_EPSILON)
#pragma unroll 1
  for (int j = 0; j < 17; ++j) {
    const int blk = j < 7 ? 0 : (j < 14 ? 1 : 2);
    double col = 0.0;
    if (free_mask & (1u << blk)) {
      const double xj = x[j];
      const double delta = fmax(min_step, fabs(xj) * 1e-6);
      x[j] = __dadd_rn(xj, delta);
      const double fp = text_residual_exact(x, rx, ry, im, fx, fy, cx, cy, mu, sigma, iref, wT);
      x[j] = __dsub_rn(xj, delta);
      const double fm = text_residual_exact(x, rx, ry, im, fx, fy, cx, cy, mu, sigma, iref, wT);
      x[j] = xj;
      col = __dmul_rn(__dsub_rn(fp, fm), __ddiv_rn(__ddiv_rn(1.0, delta), 2.0));
    }
    Ja[j] = col;
  }
  // J_tangent = J_ambient * P(q)   (QuaternionParameterization::ComputeJacobian, Appendix A.1)
  {
    const double* q = c.q;
    Jr[0] = -Ja[0] * q[1] + Ja[1] * q[0] - Ja[2] * q[3] + Ja[3] * q[2];
    Jr[1] = -Ja[0] * q[2] + Ja[1] * q[3] + Ja[2] * q[0] - Ja[3] * q[1];
    Jr[2] = -Ja[0] * q[3] - Ja[1] * q[2] + Ja[2] * q[1] + Ja[3] * q[0];
    Jr[3] = Ja[4]; Jr[4] = Ja[5]; Jr[5] = Ja[6];
  }
  {
    const double* q = h.q; const double* a = Ja + 7;
    Jr[6] = -a[0] * q[1] + a[1] * q[0] - a[2] * q[3] + a[3] * q[2];
    Jr[7] = -a[0] * q[2] + a[1] * q[3] + a[2] * q[0] - a[3] * q[1];
    Jr[8] = -a[0] * q[3] - a[1] * q[2] + a[2] * q[1] + a[3] * q[0];
    Jr[9] = Ja[11]; Jr[10] = Ja[12]; Jr[11] = Ja[13];
  }
  Jr[12] = Ja[14]; Jr[13] = Ja[15]; Jr[14] = Ja[16];
  return res;
}

// ceres::HuberLoss(a) on s: returns sqrt(rho') (the residual/Jacobian scaling) and rho0 in *cost2.
__device__ __forceinline__ double huber_scale(double a, double s, double* rho0) {
  const double b = a * a;
  if (a <= 0.0 || s <= b) { *rho0 = s; return 1.0; }
  const double r = sqrt(s);
  *rho0 = 2.0 * a * r - b;
  return sqrt(fmax(DBL_MIN, a / r));
}

}  // namespace tsl
