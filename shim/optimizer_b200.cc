// TextSLAM::optimizer on libtslam_b200.so — see optimizer_b200.h. Every Pyr* function below walks the object graph exactly where
// the reference's does (the cited line ranges), but instead of heap-allocating one ceres::CostFunction per observation it
// appends one row to a flat observation-major problem (tslam_ba_problem); matrices the reference hands to functors as
// constants (Trw / Twr / Tcr) become camera entries with cam_fixed = 1, landmarks hosted outside the window become entries with
// rho_fixed / theta_fixed = 1 — the functor zoo is the same geometry with different blocks held constant.
#include "optimizer_b200.h"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>

namespace TextSLAM {

namespace {

void die(const char* what) {   // the reference's error convention: cerr + exit(-1) (src/optimizer.cc:881-884, 1842-1845)
  std::fprintf(stderr, "optimizer(b200): %s: %s\n", what, tslam_last_error());
  std::exit(-1);
}

// Eigen::Quaterniond(R).normalized() -> (w, x, y, z)
void rot_to_quat(const Mat33& R, double q[4]) {
  const double tr = R(0, 0) + R(1, 1) + R(2, 2);
  if (tr > 0.0) {
    double t = std::sqrt(tr + 1.0);
    q[0] = 0.5 * t; t = 0.5 / t;
    q[1] = (R(2, 1) - R(1, 2)) * t; q[2] = (R(0, 2) - R(2, 0)) * t; q[3] = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    q[1 + i] = 0.5 * t; t = 0.5 / t;
    q[0] = (R(k, j) - R(j, k)) * t; q[1 + j] = (R(j, i) + R(i, j)) * t; q[1 + k] = (R(k, i) + R(i, k)) * t;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int c = 0; c < 4; ++c) q[c] /= n;
}
void pose_of(const Mat33& Rcw, const Mat31& tcw, double pose[7]) { rot_to_quat(Rcw, pose); pose[4] = tcw(0); pose[5] = tcw(1); pose[6] = tcw(2); }
void pose_of(const Mat44& T, double pose[7]) {
  Mat33 R; Mat31 t;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R(i, j) = T(i, j); t(i) = T(i, 3); }
  pose_of(R, t, pose);
}
void quat_to_rot(const double q_[4], double R[9]) {   // normalized().toRotationMatrix()
  const double n = std::sqrt(q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]);
  const double w = q_[0] / n, x = q_[1] / n, y = q_[2] / n, z = q_[3] / n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
Mat44 mat_of_pose(const double pose[7]) {   // tool::Pose2Mat44
  double R[9];
  quat_to_rot(pose, R);
  Mat44 T; T.setIdentity();
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T(i, j) = R[3 * i + j]; T(i, 3) = pose[4 + i]; }
  return T;
}
// tool::GetProjText (src/tool.cc:1593-1738): box-corner ray of the host keyframe through plane theta into the observing image.
void proj_text(const Vec2& ray, const double theta[3], const double cam[7], const double host[7], const Mat33& Kl, double uv[2]) {
  double Rc[9], Rh[9], R[9], t[3];
  quat_to_rot(cam, Rc); quat_to_rot(host, Rh);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = Rc[3 * i] * Rh[3 * j] + Rc[3 * i + 1] * Rh[3 * j + 1] + Rc[3 * i + 2] * Rh[3 * j + 2];
  for (int i = 0; i < 3; ++i) t[i] = cam[4 + i] - (R[3 * i] * host[4] + R[3 * i + 1] * host[5] + R[3 * i + 2] * host[6]);
  const double rx = ray(0), ry = ray(1);
  const double rho = -(rx * theta[0] + ry * theta[1] + theta[2]);   // TextProj, include/ModelTool.hpp:164-171
  const double px = (R[0] * rx + R[1] * ry + R[2]) / rho + t[0], py = (R[3] * rx + R[4] * ry + R[5]) / rho + t[1], pz = (R[6] * rx + R[7] * ry + R[8]) / rho + t[2];
  uv[0] = Kl(0, 0) * px / pz + Kl(0, 2);
  uv[1] = Kl(1, 1) * py / pz + Kl(1, 2);
}

}  // namespace

// Flat problem under construction: the counterpart of the ceres::Problem a Pyr* function builds.
struct optimizer::Flat {
  std::vector<double> cams, rho, theta, p_uv, p_ray, t_rays, t_iref, t_musigma;
  std::vector<uint8_t> cam_fixed, rho_fixed, theta_fixed, imgs;
  std::vector<int32_t> p_cam, p_host, p_lm, t_cam, t_host, t_plane, t_img;
  int img_w = 0, img_h = 0, n_imgs = 0;
  std::map<const keyframe*, int> const_cam, img_of;
  // text objects of this level: their quads (for mu / sigma) and block ranges
  std::vector<double> quads; std::vector<int32_t> quad_img; std::vector<int> quad_first_block, quad_nblocks;

  int add_cam(const double pose[7], bool fixed) { for (int c = 0; c < 7; ++c) cams.push_back(pose[c]); cam_fixed.push_back(fixed ? 1 : 0); return (int)cam_fixed.size() - 1; }
  int constant_cam(const keyframe* kf) {   // Trw / Twr of a keyframe outside the optimised set: one constant camera per keyframe
    auto it = const_cam.find(kf);
    if (it != const_cam.end()) return it->second;
    double pose[7];
    pose_of(kf->mTcw, pose);
    return const_cam[kf] = add_cam(pose, true);
  }
  int add_rho(double r, bool fixed) { rho.push_back(r); rho_fixed.push_back(fixed ? 1 : 0); return (int)rho.size() - 1; }
  int add_theta(const double t[3], bool fixed) { for (int c = 0; c < 3; ++c) theta.push_back(t[c]); theta_fixed.push_back(fixed ? 1 : 0); return (int)theta_fixed.size() - 1; }
  void add_point(const Vec2& uv, const Vec3& ray, int cam, int host, int lm) {
    p_uv.push_back(uv(0)); p_uv.push_back(uv(1)); p_ray.push_back(ray(0)); p_ray.push_back(ray(1));
    p_cam.push_back(cam); p_host.push_back(host); p_lm.push_back(lm);
  }
  int image(const keyframe* kf, const cv::Mat& im) {
    auto it = img_of.find(kf);
    if (it != img_of.end()) return it->second;
    if (n_imgs == 0) { img_w = im.cols; img_h = im.rows; }
    if (im.cols != img_w || im.rows != img_h) throw std::runtime_error("pyramid images of one level differ in size");
    for (int r = 0; r < im.rows; ++r) imgs.insert(imgs.end(), im.data + (size_t)r * im.step, im.data + (size_t)r * im.step + im.cols);
    return img_of[kf] = n_imgs++;
  }
  void begin_text_object(const double quad[8], int img) {
    for (int c = 0; c < 8; ++c) quads.push_back(quad[c]);
    quad_img.push_back(img); quad_first_block.push_back((int)t_cam.size()); quad_nblocks.push_back(0);
  }
  void add_text_block(const TextFeature* f, int cam, int host, int plane, int img) {
    for (int k = 0; k < 8; ++k) { t_rays.push_back(f->neighbourRay[k](0)); t_rays.push_back(f->neighbourRay[k](1)); t_iref.push_back(f->neighbourNInten[k]); }
    t_musigma.push_back(0.0); t_musigma.push_back(0.0);
    t_cam.push_back(cam); t_host.push_back(host); t_plane.push_back(plane); t_img.push_back(img);
    quad_nblocks.back()++;
  }
  // tool::CalTextinfo for every text object of the level (src/tool.cc:1178-1262), one batched call
  void fill_musigma(tslam_ctx* ctx) {
    const int nq = (int)quad_img.size();
    if (nq == 0) return;
    std::vector<double> mu(nq), sg(nq); std::vector<int32_t> ok(nq);
    if (tslam_text_info(ctx, imgs.data(), n_imgs, img_w, img_h, quads.data(), quad_img.data(), nq, mu.data(), sg.data(), ok.data())) die("tslam_text_info");
    for (int q = 0; q < nq; ++q)
      for (int b = quad_first_block[q]; b < quad_first_block[q] + quad_nblocks[q]; ++b) { t_musigma[2 * b] = mu[q]; t_musigma[2 * b + 1] = sg[q]; }
  }
  tslam_ba_problem view(const Mat33& K0, double ws, double huber_s, const Mat33& Kl, double wt, double huber_t) {
    tslam_ba_problem p;
    std::memset(&p, 0, sizeof(p));
    p.n_cams = (int)cam_fixed.size(); p.cams = cams.data(); p.cam_fixed = cam_fixed.data();
    p.n_points = (int)rho.size(); p.rho = rho.data(); p.rho_fixed = rho_fixed.data();
    p.n_planes = (int)theta_fixed.size(); p.theta = theta.data(); p.theta_fixed = theta_fixed.data();
    p.n_pobs = (int)p_cam.size(); p.p_uv = p_uv.data(); p.p_ray = p_ray.data(); p.p_cam = p_cam.data(); p.p_host = p_host.data(); p.p_lm = p_lm.data();
    p.K_point[0] = K0(0, 0); p.K_point[1] = K0(1, 1); p.K_point[2] = K0(0, 2); p.K_point[3] = K0(1, 2);
    p.w_point[0] = p.w_point[1] = ws; p.huber_point = huber_s;
    p.n_tobs = (int)t_cam.size(); p.t_rays = t_rays.data(); p.t_iref = t_iref.data(); p.t_musigma = t_musigma.data();
    p.t_cam = t_cam.data(); p.t_host = t_host.data(); p.t_plane = t_plane.data(); p.t_img = t_img.data();
    p.n_imgs = n_imgs; p.img_w = img_w; p.img_h = img_h; p.imgs = imgs.data();
    p.K_text[0] = Kl(0, 0); p.K_text[1] = Kl(1, 1); p.K_text[2] = Kl(0, 2); p.K_text[3] = Kl(1, 2);
    p.w_text = wt; p.huber_text = huber_t;
    return p;
  }
};

optimizer::optimizer(Mat33& mK, double& dScale, int& nLevels, bool& Flag_noText, bool& Flag_rapid) : K(mK), bFlag_noText(Flag_noText), bFlag_rapid(Flag_rapid) {
  vK.resize(nLevels);   // src/optimizer.cc:30-52
  vK[0] = K;
  const double invScale = 1.0 / dScale;
  for (int i = 1; i < nLevels; ++i) {
    vK[i] = vK[i - 1];
    for (double& v : vK[i].m) v *= invScale;
    vK[i](2, 2) = 1.0;
  }
  std::memset(&last_summary, 0, sizeof(last_summary));
  if (tslam_ctx_create(0, &ctx)) die("tslam_ctx_create");   // no CPU fallback: without an sm_100 device this is fatal, like a failed Ceres solve
}
optimizer::~optimizer() { if (ctx) tslam_ctx_destroy(ctx); }

static tslam_solve_options solve_opts(int its, int jac_mode) {
  tslam_solve_options o;
  std::memset(&o, 0, sizeof(o));
  o.max_iters = its; o.text_jac_mode = jac_mode; o.n_threads = 1;
  return o;
}
static tslam_gate_options gate_opts(bool scene, bool text, double ws, double chi2Mono, double wt, double chi2Text) {
  tslam_gate_options g;
  std::memset(&g, 0, sizeof(g));
  g.gate_points = scene; g.gate_text = text; g.w_point[0] = g.w_point[1] = ws; g.chi2_mono = chi2Mono;
  g.relax_below_text_blocks = 50; g.relax_amount = 4.0; g.w_text = wt; g.chi2_text = chi2Text; g.text_ratio = 0.99;
  return g;
}

// ---------------------------------------------------------------------------------------------------------------------
// PoseOptim (src/optimizer.cc:135-195) -> PyrPoseOptim (:1060-1327)
// ---------------------------------------------------------------------------------------------------------------------
void optimizer::PoseOptim(frame& F) {
  std::vector<TextObservation*> TextObjs;
  std::vector<int> FLAGTextObjs;
  std::vector<bool> vTextsGood;
  std::vector<std::vector<bool>> vTextFeatsGood;
  for (size_t i0 = 0; i0 < F.vObvText.size(); i0++) {
    if (!(F.vObvText[i0]->obj->STATE == TEXTGOOD)) continue;
    FLAGTextObjs.push_back((int)i0);
    TextObjs.push_back(F.vObvText[i0]);
    vTextsGood.push_back(F.vObvGoodTexts[i0]);
    vTextFeatsGood.push_back(F.vObvGoodTextFeats[i0]);
  }
  double pose[7];
  pose_of(F.mRcw, F.mtcw, pose);
  const double chi2Mono[4] = {12.25, 12.25, 12.25, 12.25}, chi2Text[4] = {0.5, 0.5, 0.5, 0.95};
  const int its[4] = {10, 10, 10, 10};
  std::vector<bool> vPtsGood = F.vObvGoodPts;
  for (int lv = bFlag_rapid ? 0 : 1; lv < 4; ++lv) {   // PyBegin 3 (rapid only), 2, 1, 0
    PyrPoseOptim(F, pose, 3 - lv, chi2Mono[lv], chi2Text[lv], its[lv], vPtsGood, vTextsGood, vTextFeatsGood, TextObjs);
    F.vObvGoodPts = vPtsGood;   // :1311-1317
    for (size_t k = 0; k < FLAGTextObjs.size(); ++k) { F.vObvGoodTexts[FLAGTextObjs[k]] = vTextsGood[k]; F.vObvGoodTextFeats[FLAGTextObjs[k]] = vTextFeatsGood[k]; }
  }
  F.SetPose(mat_of_pose(pose));
}

void optimizer::PyrPoseOptim(frame& F, double* pose, int PyBegin, double chi2Mono, double chi2Text, int its, std::vector<bool>& vPtsGood,
                             std::vector<bool>& vTextsGood, std::vector<std::vector<bool>>& vTextFeatsGood, const std::vector<TextObservation*>& TextObjs) {
  const double weight_S = 1.0 / 1.2, weight_T = 1.0 / 0.2;   // :1087-1088
  Flat P;
  const int cam = P.add_cam(pose, false);
  std::vector<int> vIdx2vPtsGood, vIdx2vTextsGood, vIdx2vTextFeatsGood;
  // A) scene points: auto_PoseOptimScene = the BA geometry with host pose and inverse depth constant (:1121-1147)
  const std::vector<SceneFeature*>& vObv = F.vSceneObv2d[PyBegin];
  for (size_t i0 = 0; i0 < vObv.size(); i0++) {
    const int Idx3d = vObv[i0]->IdxToRaw;
    if (!vPtsGood[(size_t)Idx3d]) continue;
    mapPts* pt = F.vObvPts[Idx3d]->pt;
    const Vec3 rayrho = pt->GetPtInv();
    P.add_point(F.vSceneObv2d[0][Idx3d]->feature, rayrho, cam, P.constant_cam(pt->RefKF), P.add_rho(rayrho(2), true));
    vIdx2vPtsGood.push_back(Idx3d);
  }
  // B) text objects: nume_PoseOptimText (:1160-1206); mu / sigma from the box projected with the current pose (:1179-1184)
  std::vector<int32_t> t_obj, obj_size(TextObjs.size(), 0);
  if (!bFlag_noText) {
    for (size_t itext = 0; itext < TextObjs.size(); itext++) {
      if (!vTextsGood[itext]) continue;
      mapText* obj = TextObjs[itext]->obj;
      const Mat31 thetaM = obj->RefKF->mNcr[obj->GetNidx()];
      const double th[3] = {thetaM(0), thetaM(1), thetaM(2)};
      const int host = P.constant_cam(obj->RefKF), plane = P.add_theta(th, true), img = P.image(&F, F.vFrameImg[PyBegin]);
      double quad[8];
      for (size_t iBox = 0; iBox < 4 && iBox < obj->vTextDeteRay.size(); iBox++) proj_text(obj->vTextDeteRay[iBox], th, pose, &P.cams[7 * host], vK[PyBegin], quad + 2 * iBox);
      P.begin_text_object(quad, img);
      const std::vector<TextFeature*>& refRay = obj->vRefFeature[PyBegin];
      for (size_t ifeat = 0; ifeat < refRay.size(); ifeat++) {
        if (!vTextFeatsGood[itext][refRay[ifeat]->IdxToRaw]) continue;
        P.add_text_block(refRay[ifeat], cam, host, plane, img);
        t_obj.push_back((int32_t)itext); obj_size[itext]++;
        vIdx2vTextsGood.push_back((int)itext); vIdx2vTextFeatsGood.push_back(refRay[ifeat]->IdxToRaw);
      }
    }
  }
  P.fill_musigma(ctx);
  tslam_ba_problem prob = P.view(vK[0], weight_S, std::sqrt(5.991), vK[PyBegin], weight_T, 3.0);
  const tslam_solve_options so = solve_opts(its, text_jac_mode);
  const tslam_gate_options go = gate_opts(!bFlag_rapid, !bFlag_rapid, weight_S, chi2Mono, weight_T, chi2Text);   // SCENEOutlier / TEXTOutlier (:1077-1080)
  std::vector<uint8_t> pt_bad(prob.n_pobs + 1), tf_bad(prob.n_tobs + 1), obj_bad(TextObjs.size() + 1);
  if (tslam_solve_gated(ctx, &prob, &so, &go, t_obj.data(), obj_size.data(), (int)TextObjs.size(), &last_summary, nullptr, nullptr, pt_bad.data(), tf_bad.data(),
                        obj_bad.data(), nullptr))
    die("PyrPoseOptim");
  for (int c = 0; c < 7; ++c) pose[c] = P.cams[7 * cam + c];
  // outlier flags (:1236-1302)
  for (int i = 0; i < prob.n_pobs; ++i) if (pt_bad[i]) vPtsGood[vIdx2vPtsGood[i]] = false;
  for (int j = 0; j < prob.n_tobs; ++j) if (tf_bad[j]) vTextFeatsGood[vIdx2vTextsGood[j]][vIdx2vTextFeatsGood[j]] = false;
  for (size_t o = 0; o < TextObjs.size(); ++o) if (obj_bad[o]) vTextsGood[o] = false;
}

// ---------------------------------------------------------------------------------------------------------------------
// LocalBundleAdjustment (src/optimizer.cc:197-331) -> PyrBA (:1330-1698)
// ---------------------------------------------------------------------------------------------------------------------
static void flatten_map(map* mpMap, const std::vector<keyframe*>& vKFs, std::vector<mapPts*>& vMapPts, std::vector<mapText*>& vMapTexts, std::vector<int>& vmnId2MapPts,
                        std::vector<int>& vmnId2MapTexts, std::vector<int>& vmnId2vKFs, double**& rho, double**& theta, double**& pose) {
  vMapPts = mpMap->GetAllMapPoints();
  vMapTexts = mpMap->GetAllMapTexts(TEXTGOOD);
  vmnId2MapPts.assign(mpMap->imapPts, -1); vmnId2MapTexts.assign(mpMap->imapText, -1); vmnId2vKFs.assign(mpMap->imapkfs, -1);
  rho = new double*[vMapPts.size() + 1]; theta = new double*[vMapTexts.size() + 1]; pose = new double*[vKFs.size() + 1];
  for (size_t i = 0; i < vKFs.size(); ++i) { vmnId2vKFs[vKFs[i]->mnId] = (int)i; pose[i] = new double[7]; pose_of(vKFs[i]->mRcw, vKFs[i]->mtcw, pose[i]); }
  for (size_t i = 0; i < vMapPts.size(); ++i) { rho[i] = new double[1]; rho[i][0] = vMapPts[i]->GetInverD(); vmnId2MapPts[vMapPts[i]->mnId] = (int)i; }
  for (size_t i = 0; i < vMapTexts.size(); ++i) {
    theta[i] = new double[3];
    const Mat31 n = vMapTexts[i]->RefKF->mNcr[(size_t)vMapTexts[i]->GetNidx()];
    theta[i][0] = n(0); theta[i][1] = n(1); theta[i][2] = n(2);
    vmnId2MapTexts[vMapTexts[i]->mnId] = (int)i;
  }
}
static void free_flat(double** a, size_t n) { for (size_t i = 0; i < n; ++i) delete[] a[i]; delete[] a; }   // (the reference leaks these, :213-225)

void optimizer::LocalBundleAdjustment(map* mpMap, std::vector<keyframe*> vKFs, const BAStatus& STATE) {
  std::vector<mapPts*> vMapPts; std::vector<mapText*> vMapTexts;
  std::vector<int> vmnId2MapPts, vmnId2MapTexts, vmnId2vKFs;
  double **rho, **theta, **pose;
  flatten_map(mpMap, vKFs, vMapPts, vMapTexts, vmnId2MapPts, vmnId2MapTexts, vmnId2vKFs, rho, theta, pose);
  std::vector<bool> vMapPtOptim(vMapPts.size()), vMapTextOptim(vMapTexts.size());
  for (size_t i = 0; i < vMapPts.size(); ++i) vMapPtOptim[i] = vmnId2vKFs[vMapPts[i]->RefKF->mnId] >= 0;       // landmark hosted outside the window -> constant (:240-243)
  for (size_t i = 0; i < vMapTexts.size(); ++i) vMapTextOptim[i] = vmnId2vKFs[vMapTexts[i]->RefKF->mnId] >= 0;
  std::vector<int> InitialIdx;
  for (size_t i = 0; i < vKFs.size(); ++i) if (vKFs[i]->mnId == 0 || vKFs[i]->mnId == 1) InitialIdx.push_back((int)i);   // :277-278
  const double chi2Mono[4] = {12.25, 12.25, 12.25, 12.25}, chi2Text[4] = {0.5, 0.5, 0.5, 0.95};
  const int its[4] = {10, 10, 10, 10};
  for (int lv = 1; lv < 4; ++lv)   // PyBegin 2, 1, 0 (:282-289)
    PyrBA(pose, theta, rho, vKFs, vmnId2MapPts, vMapPtOptim, vmnId2MapTexts, vMapTextOptim, vmnId2vKFs, InitialIdx, 3 - lv, chi2Mono[lv], chi2Text[lv], its[lv], STATE);
  for (size_t i = 0; i < vKFs.size(); ++i) vKFs[i]->SetPose(mat_of_pose(pose[i]));   // :294-326
  for (size_t i = 0; i < vMapPts.size(); ++i) vMapPts[i]->SetRho(rho[i][0]);
  for (size_t i = 0; i < vMapTexts.size(); ++i) { Mat31 n; n(0) = theta[i][0]; n(1) = theta[i][1]; n(2) = theta[i][2]; vMapTexts[i]->RefKF->SetN(n, vMapTexts[i]->GetNidx()); }
  free_flat(rho, vMapPts.size()); free_flat(theta, vMapTexts.size()); free_flat(pose, vKFs.size());
}

void optimizer::PyrBA(double** pose, double** theta, double** rho, const std::vector<keyframe*>& vKFs, const std::vector<int>& vmnId2Pts, const std::vector<bool>& vPtOptim,
                      const std::vector<int>& vmnId2Texts, const std::vector<bool>& vTextOptim, const std::vector<int>& vmnId2vKFs, const std::vector<int>& InitialIdx,
                      int PyBegin, double chi2Mono, double chi2Text, int its, const BAStatus& STATE) {
  const double weight_S = 1.0 / 1.2, weight_T = 1.0 / 0.2;   // :1350-1351
  Flat P;
  for (size_t iKF = 0; iKF < vKFs.size(); iKF++) P.add_cam(pose[iKF], false);   // camera k == keyframe k of the window; constants are appended behind
  std::vector<int> rho_of(vmnId2Pts.size(), -1), theta_of(vmnId2Texts.size(), -1);   // map index -> entry of this level's problem
  std::vector<bool> FLAG_KFIN(vKFs.size(), false);
  std::vector<int> vIdx2vPtsGood, vIdxS2vKFs, vIdxT2vKFs, vIdx2Texts, vIdx2TextFeats;
  // A) scene points (:1365-1436)
  for (size_t iKF = 0; iKF < vKFs.size(); iKF++) {
    keyframe* kf = vKFs[iKF];
    const std::vector<SceneFeature*>& vSceneObv = kf->vSceneObv2d[PyBegin];
    for (size_t iScene = 0; iScene < vSceneObv.size(); iScene++) {
      const int Idx2Raw = vSceneObv[iScene]->IdxToRaw;
      if (!kf->vObvGoodPts[Idx2Raw]) continue;
      mapPts* pt = kf->vObvPts[(size_t)Idx2Raw]->pt;
      const int IdxRho = vmnId2Pts[pt->mnId], IdxRef = vmnId2vKFs[pt->RefKF->mnId];
      int host, lm;
      if (vPtOptim[IdxRho]) {                 // auto_BAScene (:1394-1418)
        if (IdxRef == (int)iKF) continue;     // host != target
        if (rho_of[IdxRho] < 0) rho_of[IdxRho] = P.add_rho(rho[IdxRho][0], false);
        host = IdxRef; lm = rho_of[IdxRho];
        FLAG_KFIN[IdxRef] = true;
      } else {                                // auto_PoseOptimScene: host pose and inverse depth constant (:1419-1430)
        host = P.constant_cam(pt->RefKF); lm = P.add_rho(pt->GetPtInv()(2), true);
      }
      P.add_point(kf->vSceneObv2d[0][Idx2Raw]->feature, pt->GetRaydir(), (int)iKF, host, lm);
      FLAG_KFIN[iKF] = true;
      vIdxS2vKFs.push_back((int)iKF); vIdx2vPtsGood.push_back(Idx2Raw);
    }
  }
  // B) text objects (:1447-1557)
  std::vector<int32_t> t_obj, obj_size;
  std::vector<std::pair<int, int>> obj_kf_raw;   // (keyframe, raw object index) of every counted object (idx_texts)
  if (!bFlag_noText) {
    for (size_t iKF = 0; iKF < vKFs.size(); iKF++) {
      keyframe* kf = vKFs[iKF];
      std::vector<int> vNew2RawTextkf;
      const std::vector<TextObservation*> vText = kf->GetStateTextObvs(TEXTGOOD, vNew2RawTextkf);
      for (size_t iobj = 0; iobj < vText.size(); iobj++) {
        obj_size.push_back(0);
        const int idx_texts = (int)obj_size.size() - 1, idxRawObj = vNew2RawTextkf[iobj];
        obj_kf_raw.emplace_back((int)iKF, idxRawObj);
        if (!kf->vObvGoodTexts[idxRawObj]) continue;
        mapText* obj = vText[iobj]->obj;
        const int idxtheta = vmnId2Texts[obj->mnId], idxref = vmnId2vKFs[obj->RefKF->mnId];
        int host, plane;
        double th[3];
        if (vTextOptim[idxtheta]) {             // nume_BAText (:1482-1522)
          if (idxref == (int)iKF) continue;
          if (theta_of[idxtheta] < 0) theta_of[idxtheta] = P.add_theta(theta[idxtheta], false);
          host = idxref; plane = theta_of[idxtheta];
          for (int c = 0; c < 3; ++c) th[c] = theta[idxtheta][c];
        } else {                                // nume_PoseOptimText: plane and host pose constant (:1523-1554)
          const Mat31 n = obj->RefKF->mNcr[obj->GetNidx()];
          th[0] = n(0); th[1] = n(1); th[2] = n(2);
          host = P.constant_cam(obj->RefKF); plane = P.add_theta(th, true);
        }
        const int img = P.image(kf, kf->vFrameImg[PyBegin]);
        double quad[8];
        for (size_t iBox = 0; iBox < 4 && iBox < obj->vTextDeteRay.size(); iBox++) proj_text(obj->vTextDeteRay[iBox], th, pose[iKF], &P.cams[7 * host], vK[PyBegin], quad + 2 * iBox);
        P.begin_text_object(quad, img);
        const std::vector<TextFeature*>& refRay = obj->vRefFeature[PyBegin];
        for (size_t ifeat = 0; ifeat < refRay.size(); ifeat++) {
          if (!kf->vObvGoodTextFeats[idxRawObj][refRay[ifeat]->IdxToRaw]) continue;
          P.add_text_block(refRay[ifeat], (int)iKF, host, plane, img);
          t_obj.push_back(idx_texts); obj_size[idx_texts]++;
          vIdxT2vKFs.push_back((int)iKF); vIdx2Texts.push_back(idxRawObj); vIdx2TextFeats.push_back(refRay[ifeat]->IdxToRaw);
          FLAG_KFIN[iKF] = true;
          if (vTextOptim[idxtheta]) FLAG_KFIN[idxref] = true;
        }
      }
    }
  }
  // fixed keyframes (:1562-1588): the first two keyframes of the map, and in LOCAL state the first three participating ones
  for (int k : InitialIdx) if (FLAG_KFIN[k]) P.cam_fixed[k] = 1;
  if (STATE == LOCAL) {
    int num = 0, fixed = 0;
    for (bool f : FLAG_KFIN) num += f;
    if (num > 3)
      for (size_t k = 0; k < FLAG_KFIN.size() && fixed < 3; ++k) if (FLAG_KFIN[k]) { P.cam_fixed[k] = 1; ++fixed; }
  }
  P.fill_musigma(ctx);
  tslam_ba_problem prob = P.view(vK[0], weight_S, std::sqrt(5.991), vK[PyBegin], weight_T, 3.0);
  const tslam_solve_options so = solve_opts(its, text_jac_mode);
  const tslam_gate_options go = gate_opts(!bFlag_rapid, !bFlag_rapid, weight_S, chi2Mono, weight_T, chi2Text);
  std::vector<uint8_t> pt_bad(prob.n_pobs + 1), tf_bad(prob.n_tobs + 1), obj_bad(obj_size.size() + 1);
  if (tslam_solve_gated(ctx, &prob, &so, &go, t_obj.data(), obj_size.data(), (int)obj_size.size(), &last_summary, nullptr, nullptr, pt_bad.data(), tf_bad.data(),
                        obj_bad.data(), nullptr))
    die("PyrBA");
  // parameters back into the caller's blocks (Ceres optimises them in place)
  for (size_t k = 0; k < vKFs.size(); ++k) for (int c = 0; c < 7; ++c) pose[k][c] = P.cams[7 * k + c];
  for (size_t l = 0; l < rho_of.size(); ++l) if (rho_of[l] >= 0) rho[l][0] = P.rho[rho_of[l]];
  for (size_t t = 0; t < theta_of.size(); ++t) if (theta_of[t] >= 0) for (int c = 0; c < 3; ++c) theta[t][c] = P.theta[3 * theta_of[t] + c];
  // outlier flags (:1616-1684)
  for (int i = 0; i < prob.n_pobs; ++i) if (pt_bad[i]) vKFs[vIdxS2vKFs[i]]->vObvGoodPts[vIdx2vPtsGood[i]] = false;
  for (int j = 0; j < prob.n_tobs; ++j) if (tf_bad[j]) vKFs[vIdxT2vKFs[j]]->vObvGoodTextFeats[vIdx2Texts[j]][vIdx2TextFeats[j]] = false;
  for (size_t o = 0; o < obj_size.size(); ++o) if (obj_bad[o]) vKFs[obj_kf_raw[o].first]->vObvGoodTexts[obj_kf_raw[o].second] = false;
}

// ---------------------------------------------------------------------------------------------------------------------
// GlobalBA (src/optimizer.cc:334-453) -> PyrGlobalBA (:1701-1851): level 0, 20 iterations, unweighted points, text branch off
// ---------------------------------------------------------------------------------------------------------------------
void optimizer::GlobalBA(map* mpMap) {
  std::vector<keyframe*> vKFs = mpMap->GetAllKeyFrame();
  std::vector<mapPts*> vMapPts; std::vector<mapText*> vMapTexts;
  std::vector<int> vmnId2MapPts, vmnId2MapTexts, vmnId2vKFs;
  double **rho, **theta, **pose;
  flatten_map(mpMap, vKFs, vMapPts, vMapTexts, vmnId2MapPts, vmnId2MapTexts, vmnId2vKFs, rho, theta, pose);
  std::vector<int> InitialIdx;
  for (size_t i = 0; i < vKFs.size(); ++i) if (vKFs[i]->mnId == 0 || vKFs[i]->mnId == 1) InitialIdx.push_back((int)i);
  PyrGlobalBA(pose, theta, rho, vKFs, vmnId2MapPts, vmnId2MapTexts, vmnId2vKFs, InitialIdx, 0, 20);   // :411-414
  for (size_t i = 0; i < vKFs.size(); ++i) vKFs[i]->SetPose(mat_of_pose(pose[i]));
  for (size_t i = 0; i < vMapPts.size(); ++i) vMapPts[i]->SetRho(rho[i][0]);
  for (size_t i = 0; i < vMapTexts.size(); ++i) { Mat31 n; n(0) = theta[i][0]; n(1) = theta[i][1]; n(2) = theta[i][2]; vMapTexts[i]->RefKF->SetN(n, vMapTexts[i]->GetNidx()); }
  free_flat(rho, vMapPts.size()); free_flat(theta, vMapTexts.size()); free_flat(pose, vKFs.size());
}

void optimizer::PyrGlobalBA(double** pose, double** theta, double** rho, const std::vector<keyframe*>& vKFs, const std::vector<int>& vmnId2Pts,
                            const std::vector<int>& vmnId2Texts, const std::vector<int>& vmnId2vKFs, const std::vector<int>& InitialIdx, int PyBegin, int its) {
  (void)theta; (void)vmnId2Texts;   // FLAG_TEXT = false (:1707): the text branch (:1766-1822) is dead code in the reference
  Flat P;
  for (size_t iKF = 0; iKF < vKFs.size(); iKF++) P.add_cam(pose[iKF], false);
  std::vector<int> rho_of(vmnId2Pts.size(), -1);
  std::vector<bool> FLAG_KFIN(vKFs.size(), false);
  for (size_t iKF = 0; iKF < vKFs.size(); iKF++) {   // :1716-1760 (the Good flags are NOT consulted here)
    keyframe* kf = vKFs[iKF];
    const std::vector<SceneFeature*>& vSceneObv = kf->vSceneObv2d[PyBegin];
    for (size_t iScene = 0; iScene < vSceneObv.size(); iScene++) {
      const int Idx2Raw = vSceneObv[iScene]->IdxToRaw;
      mapPts* pt = kf->vObvPts[(size_t)Idx2Raw]->pt;
      const int IdxRho = vmnId2Pts[pt->mnId], IdxRef = vmnId2vKFs[pt->RefKF->mnId];
      if (IdxRho < 0 || IdxRef == (int)iKF) continue;
      if (rho_of[IdxRho] < 0) rho_of[IdxRho] = P.add_rho(rho[IdxRho][0], false);
      P.add_point(kf->vSceneObv2d[0][Idx2Raw]->feature, pt->GetRaydir(), (int)iKF, IdxRef, rho_of[IdxRho]);   // auto_BASceneNW: unweighted
      FLAG_KFIN[iKF] = true; FLAG_KFIN[IdxRef] = true;
    }
  }
  for (int k : InitialIdx) if (FLAG_KFIN[k]) P.cam_fixed[k] = 1;   // :1825-1829
  tslam_ba_problem prob = P.view(vK[0], 1.0, std::sqrt(5.991), vK[PyBegin], 1.0, 3.0);
  const tslam_solve_options so = solve_opts(its, text_jac_mode);
  if (tslam_solve(ctx, &prob, &so, &last_summary, nullptr, nullptr)) die("PyrGlobalBA");
  for (size_t k = 0; k < vKFs.size(); ++k) for (int c = 0; c < 7; ++c) pose[k][c] = P.cams[7 * k + c];
  for (size_t l = 0; l < rho_of.size(); ++l) if (rho_of[l] >= 0) rho[l][0] = P.rho[rho_of[l]];
}

// ---------------------------------------------------------------------------------------------------------------------
// OptimizeLandmarker (src/optimizer.cc:456-562) -> PyrLandmarkers (:1853-2168): every pose constant
// ---------------------------------------------------------------------------------------------------------------------
void optimizer::OptimizeLandmarker(map* mpMap) {
  std::vector<keyframe*> vKFs = mpMap->GetAllKeyFrame();
  std::vector<mapPts*> vMapPts; std::vector<mapText*> vMapTexts;
  std::vector<int> vmnId2MapPts, vmnId2MapTexts, vmnId2vKFs;
  double **rho, **theta, **pose;
  flatten_map(mpMap, vKFs, vMapPts, vMapTexts, vmnId2MapPts, vmnId2MapTexts, vmnId2vKFs, rho, theta, pose);
  std::vector<std::vector<bool>> vPtsGood(vKFs.size());
  for (size_t i = 0; i < vKFs.size(); ++i) vPtsGood[i] = vKFs[i]->vObvGoodPts;
  for (int lv = 0; lv < 4; ++lv) PyrLandmarkers(pose, theta, rho, vKFs, vmnId2MapPts, vmnId2MapTexts, vPtsGood, 3 - lv, 18.0, 50);   // :532-540
  for (size_t i = 0; i < vMapPts.size(); ++i) vMapPts[i]->SetRho(rho[i][0]);
  for (size_t i = 0; i < vMapTexts.size(); ++i) { Mat31 n; n(0) = theta[i][0]; n(1) = theta[i][1]; n(2) = theta[i][2]; vMapTexts[i]->RefKF->SetN(n, vMapTexts[i]->GetNidx()); }
  free_flat(rho, vMapPts.size()); free_flat(theta, vMapTexts.size()); free_flat(pose, vKFs.size());
}

void optimizer::PyrLandmarkers(double** pose, double** theta, double** rho, const std::vector<keyframe*>& vKFs, const std::vector<int>& vmnId2Pts,
                               const std::vector<int>& vmnId2Texts, std::vector<std::vector<bool>>& vPtsGoodkf, int PyBegin, double chi2Mono, int its) {
  Flat P;
  for (size_t iKF = 0; iKF < vKFs.size(); iKF++) P.add_cam(pose[iKF], true);   // auto_RhoScene / nume_thetaText take Tcr as a constant
  std::vector<int> rho_of(vmnId2Pts.size(), -1), theta_of(vmnId2Texts.size(), -1), vIdxS2vKFs, vIdx2vPtsGood;
  std::map<const keyframe*, int> cam_of;
  for (size_t iKF = 0; iKF < vKFs.size(); iKF++) cam_of[vKFs[iKF]] = (int)iKF;
  auto cam_index = [&](const keyframe* kf) { auto it = cam_of.find(kf); return it != cam_of.end() ? it->second : P.constant_cam(kf); };
  for (size_t iKF = 0; iKF < vKFs.size(); iKF++) {   // :1869-1912
    keyframe* kf = vKFs[iKF];
    const std::vector<SceneFeature*>& vSceneObv = kf->vSceneObv2d[PyBegin];
    for (size_t iScene = 0; iScene < vSceneObv.size(); iScene++) {
      const int Idx2Raw = vSceneObv[iScene]->IdxToRaw;
      if (!vPtsGoodkf[iKF][Idx2Raw]) continue;
      mapPts* pt = kf->vObvPts[(size_t)Idx2Raw]->pt;
      const int IdxRho = vmnId2Pts[pt->mnId];
      if (IdxRho < 0 || pt->RefKF->mnId == kf->mnId) continue;
      if (rho_of[IdxRho] < 0) rho_of[IdxRho] = P.add_rho(rho[IdxRho][0], false);
      P.add_point(kf->vSceneObv2d[0][Idx2Raw]->feature, pt->GetRaydir(), (int)iKF, cam_index(pt->RefKF), rho_of[IdxRho]);   // auto_RhoScene: unweighted (include/auto_RhoScene.h:31-32)
      vIdxS2vKFs.push_back((int)iKF); vIdx2vPtsGood.push_back(Idx2Raw);
    }
  }
  std::vector<int32_t> t_obj, obj_size;
  if (!bFlag_noText) {   // :1918-1970, nume_thetaText: unweighted, HuberLoss(2.0)
    for (size_t iKF = 0; iKF < vKFs.size(); iKF++) {
      keyframe* kf = vKFs[iKF];
      std::vector<int> vNew2RawTextkf;
      const std::vector<TextObservation*> vText = kf->GetStateTextObvs(TEXTGOOD, vNew2RawTextkf);
      for (size_t iobj = 0; iobj < vText.size(); iobj++) {
        mapText* obj = vText[iobj]->obj;
        const int idxtheta = vmnId2Texts[obj->mnId];
        if (idxtheta < 0 || obj->RefKF->mnId == kf->mnId) continue;
        if (theta_of[idxtheta] < 0) theta_of[idxtheta] = P.add_theta(theta[idxtheta], false);
        const int host = cam_index(obj->RefKF), img = P.image(kf, kf->vFrameImg[PyBegin]);
        double quad[8];
        for (size_t iBox = 0; iBox < 4 && iBox < obj->vTextDeteRay.size(); iBox++) proj_text(obj->vTextDeteRay[iBox], theta[idxtheta], pose[iKF], &P.cams[7 * host], vK[PyBegin], quad + 2 * iBox);
        P.begin_text_object(quad, img);
        obj_size.push_back(0);
        for (TextFeature* f : obj->vRefFeature[PyBegin]) { P.add_text_block(f, (int)iKF, host, theta_of[idxtheta], img); t_obj.push_back((int32_t)obj_size.size() - 1); obj_size.back()++; }
      }
    }
  }
  P.fill_musigma(ctx);
  tslam_ba_problem prob = P.view(vK[0], 1.0, std::sqrt(5.991), vK[PyBegin], 1.0, 2.0);
  const tslam_solve_options so = solve_opts(its, text_jac_mode);
  tslam_gate_options g2 = gate_opts(true, false, 1.0, chi2Mono, 1.0, 1.5);   // SCENEOutlier = true, TEXTOutlier = false (:1861)
  g2.relax_below_text_blocks = 0;   // PyrLandmarkers has no "+4 when few text blocks" rule (:1992-2047)
  std::vector<uint8_t> pt_bad(prob.n_pobs + 1), tf_bad(prob.n_tobs + 1), obj_bad(obj_size.size() + 1);
  if (tslam_solve_gated(ctx, &prob, &so, &g2, t_obj.data(), obj_size.data(), (int)obj_size.size(), &last_summary, nullptr, nullptr, pt_bad.data(), tf_bad.data(),
                        obj_bad.data(), nullptr))
    die("PyrLandmarkers");
  for (size_t l = 0; l < rho_of.size(); ++l) if (rho_of[l] >= 0) rho[l][0] = P.rho[rho_of[l]];
  for (size_t t = 0; t < theta_of.size(); ++t) if (theta_of[t] >= 0) for (int c = 0; c < 3; ++c) theta[t][c] = P.theta[3 * theta_of[t] + c];
  for (int i = 0; i < prob.n_pobs; ++i) if (pt_bad[i]) vPtsGoodkf[vIdxS2vKFs[i]][vIdx2vPtsGood[i]] = false;   // :2043-2044
}

// ---------------------------------------------------------------------------------------------------------------------
// ThetaOptimMultiFs (src/optimizer.cc:565-624) -> PyrThetaOptim (:2170-2242)
// ---------------------------------------------------------------------------------------------------------------------
bool optimizer::ThetaOptimMultiFs(const frame& F, mapText*& obj) {
  const Mat31 thetaRaw = obj->RefKF->mNcr[obj->GetNidx()];
  double theta[3] = {thetaRaw(0), thetaRaw(1), thetaRaw(2)};
  Mat33 thetaVariance;
  std::vector<Mat44> vTcr;                       // every observing frame, relative to the text's host keyframe
  std::vector<std::vector<cv::Mat>> vImg(obj->RefKF->iScaleLevels);
  auto rel = [&](const Mat44& Tcw) {             // Tcw * Trw^-1
    const Mat44& Twr = obj->RefKF->mTwc;
    Mat44 T;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += Tcw(i, k) * Twr(k, j); T(i, j) = s; }
    return T;
  };
  for (auto& kv : obj->vObvkeyframe) {
    if (kv.first->mnId == obj->RefKF->mnId) continue;
    vTcr.push_back(rel(kv.first->mTcw));
    for (size_t ipy = 0; ipy < vImg.size(); ipy++) vImg[ipy].push_back(kv.first->vFrameImg[ipy]);
  }
  vTcr.push_back(rel(F.mTcw));
  for (size_t ipy = 0; ipy < vImg.size(); ipy++) vImg[ipy].push_back(F.vFrameImg[ipy]);
  for (int PyBegin = 2; PyBegin >= 0; --PyBegin)   // :601-617
    if (!PyrThetaOptim(vImg[PyBegin], vTcr, obj, PyBegin, theta, thetaVariance)) return false;
  Mat31 thetaNew; thetaNew(0) = theta[0]; thetaNew(1) = theta[1]; thetaNew(2) = theta[2];
  obj->RefKF->SetN(thetaNew, obj->GetNidx());
  obj->Covariance = thetaVariance;
  return true;   // (the reference falls off the end of the function here, :619-624)
}

bool optimizer::PyrThetaOptim(const std::vector<cv::Mat>& vImg, const std::vector<Mat44>& vTcr, mapText* obj, int PyBegin, double* theta, Mat33& thetaVariance) {
  Flat P;
  const double ident[7] = {1, 0, 0, 0, 0, 0, 0};
  const int host = P.add_cam(ident, true);       // the functor sees Tcr only: the host is the identity frame, every observer the constant Tcr
  const int plane = P.add_theta(theta, false);
  for (size_t ifs = 0; ifs < vImg.size(); ifs++) {
    double pc[7];
    pose_of(vTcr[ifs], pc);
    const int cam = P.add_cam(pc, true);
    if (P.n_imgs == 0) { P.img_w = vImg[ifs].cols; P.img_h = vImg[ifs].rows; }
    for (int r = 0; r < vImg[ifs].rows; ++r) P.imgs.insert(P.imgs.end(), vImg[ifs].data + (size_t)r * vImg[ifs].step, vImg[ifs].data + (size_t)r * vImg[ifs].step + vImg[ifs].cols);
    const int img = P.n_imgs++;
    double quad[8];
    for (size_t iBox = 0; iBox < 4 && iBox < obj->vTextDeteRay.size(); iBox++) proj_text(obj->vTextDeteRay[iBox], theta, pc, ident, vK[PyBegin], quad + 2 * iBox);
    P.begin_text_object(quad, img);
    for (TextFeature* f : obj->vRefFeature[PyBegin]) P.add_text_block(f, cam, host, plane, img);
  }
  P.fill_musigma(ctx);
  tslam_ba_problem prob = P.view(vK[0], 1.0, 0.0, vK[PyBegin], 1.0, 0.0);   // nume_thetaText: unweighted, loss_function = nullptr (:2176)
  const tslam_solve_options so = solve_opts(50, text_jac_mode);             // max_num_iterations is left at Ceres' default (:2203-2209)
  if (tslam_solve(ctx, &prob, &so, &last_summary, nullptr, nullptr)) return false;
  for (int c = 0; c < 3; ++c) theta[c] = P.theta[3 * plane + c];
  double cov[9]; int32_t n_singular = 0;                                    // ceres::Covariance of theta (:2219-2238)
  if (tslam_theta_covariance(ctx, &prob, text_jac_mode, cov, &n_singular)) return false;
  if (n_singular == 0) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) thetaVariance(i, j) = cov[3 * i + j];
  return true;
}

}  // namespace TextSLAM
