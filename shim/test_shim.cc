// Test driver of the shims: reads a flattened synthetic problem (written by tests/test_gpu_shim.py), rebuilds the TextSLAM
// object graph it came from (keyframes, inverse-depth points hosted in keyframes, text objects with their reference features,
// per-keyframe observation lists and Good flags), calls THROUGH THE CLASS SURFACE (optimizer::GlobalBA / LocalBundleAdjustment /
// PoseOptim / OptimizeLandmarker / ThetaOptimMultiFs, ORBextractor::operator()) and writes the mutated graph back for the
// Python side to compare with the flat solve of the same problem.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include "optimizer_b200.h"
#include "ORBextractor_b200.h"

using namespace TextSLAM;

template <class T>
static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }

static void quat_rot(const double* q, Mat33& R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - w * z); R(0, 2) = 2 * (x * z + w * y);
  R(1, 0) = 2 * (x * y + w * z); R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - w * x);
  R(2, 0) = 2 * (x * z - w * y); R(2, 1) = 2 * (y * z + w * x); R(2, 2) = 1 - 2 * (x * x + y * y);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: test_shim in.bin out.bin\n"); return 2; }
  if (std::string(argv[1]) == "--orb") {   // ORBextractor: raw u8 image w x h in, keypoints + descriptors out
    FILE* f = fopen(argv[2], "rb");
    int32_t wh[3];
    if (!f || fread(wh, 4, 3, f) != 3) return 2;
    std::vector<unsigned char> img = rd<unsigned char>(f, (size_t)wh[0] * wh[1]);
    fclose(f);
    ORBextractor ex(wh[2], 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kps; cv::Mat desc;
    cv::Mat im(wh[1], wh[0], img.data());
    ex(im, cv::Mat(), kps, desc);
    FILE* o = fopen(argv[3], "wb");
    int32_t n = (int32_t)kps.size();
    fwrite(&n, 4, 1, o);
    fwrite(kps.data(), sizeof(cv::KeyPoint), kps.size(), o);
    fwrite(desc.data, 1, (size_t)n * 32, o);
    fclose(o);
    return 0;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<int32_t> hd = rd<int32_t>(f, 12);
  const int n_cams = hd[0], n_points = hd[1], n_planes = hd[2], n_pobs = hd[3], n_tobs = hd[4], n_imgs = hd[5], img_w = hd[6], img_h = hd[7], mode = hd[8],
            n_window = hd[9], nlevels_i = hd[10];
  std::vector<double> K4 = rd<double>(f, 4), cams = rd<double>(f, 7 * (size_t)n_cams), rho = rd<double>(f, n_points), theta = rd<double>(f, 3 * (size_t)n_planes);
  std::vector<double> p_uv = rd<double>(f, 2 * (size_t)n_pobs), p_ray = rd<double>(f, 2 * (size_t)n_pobs);
  std::vector<int32_t> p_cam = rd<int32_t>(f, n_pobs), p_host = rd<int32_t>(f, n_pobs), p_lm = rd<int32_t>(f, n_pobs);
  std::vector<double> t_rays = rd<double>(f, 16 * (size_t)n_tobs), t_iref = rd<double>(f, 8 * (size_t)n_tobs);
  std::vector<int32_t> t_cam = rd<int32_t>(f, n_tobs), t_host = rd<int32_t>(f, n_tobs), t_plane = rd<int32_t>(f, n_tobs), t_img = rd<int32_t>(f, n_tobs);
  std::vector<double> box = rd<double>(f, 8 * (size_t)n_planes);
  std::vector<unsigned char> imgs = rd<unsigned char>(f, (size_t)n_imgs * img_w * img_h);
  fclose(f);
  const int NL = 4;
  // ---- object graph ----
  std::vector<std::unique_ptr<keyframe>> kfs(n_cams);
  std::vector<int> img_of_cam(n_cams, -1);
  for (int j = 0; j < n_tobs; ++j) img_of_cam[t_cam[j]] = t_img[j];
  for (int k = 0; k < n_cams; ++k) {
    kfs[k].reset(new keyframe());
    keyframe* kf = kfs[k].get();
    kf->mnId = k; kf->iScaleLevels = NL;
    Mat33 R; quat_rot(&cams[7 * k], R);
    Mat44 T; T.setIdentity();
    for (int i = 0; i < 3; ++i) { for (int jj = 0; jj < 3; ++jj) T(i, jj) = R(i, jj); T(i, 3) = cams[7 * k + 4 + i]; }
    kf->SetPose(T);
    kf->vSceneObv2d.resize(NL);
    kf->vFrameImg.resize(NL);
    if (img_of_cam[k] >= 0) for (int l = 0; l < NL; ++l) kf->vFrameImg[l] = cv::Mat(img_h, img_w, imgs.data() + (size_t)img_of_cam[k] * img_w * img_h);
  }
  std::vector<std::unique_ptr<mapPts>> pts(n_points);
  for (int i = 0; i < n_pobs; ++i) {
    const int l = p_lm[i];
    if (!pts[l]) { pts[l].reset(new mapPts()); pts[l]->mnId = l; pts[l]->RefKF = kfs[p_host[i]].get(); pts[l]->ray(0) = p_ray[2 * i]; pts[l]->ray(1) = p_ray[2 * i + 1]; pts[l]->ray(2) = 1.0; pts[l]->rho = rho[l]; }
  }
  std::vector<std::unique_ptr<SceneObservation>> sobs; std::vector<std::unique_ptr<SceneFeature>> sfeat;
  for (int i = 0; i < n_pobs; ++i) {
    keyframe* kf = kfs[p_cam[i]].get();
    const int idx = (int)kf->vObvPts.size();
    sobs.emplace_back(new SceneObservation()); sobs.back()->pt = pts[p_lm[i]].get();
    kf->vObvPts.push_back(sobs.back().get()); kf->vObvGoodPts.push_back(true);
    for (int l = 0; l < NL; ++l) {
      sfeat.emplace_back(new SceneFeature()); sfeat.back()->feature(0) = p_uv[2 * i]; sfeat.back()->feature(1) = p_uv[2 * i + 1]; sfeat.back()->IdxToRaw = idx;
      kf->vSceneObv2d[l].push_back(sfeat.back().get());
    }
  }
  std::vector<std::unique_ptr<mapText>> texts(n_planes);
  std::vector<std::unique_ptr<TextFeature>> tfeat; std::vector<std::unique_ptr<TextObservation>> tobs;
  for (int j = 0; j < n_tobs;) {   // blocks of one (keyframe, plane) are contiguous
    int e = j;
    while (e < n_tobs && t_cam[e] == t_cam[j] && t_plane[e] == t_plane[j]) ++e;
    const int t = t_plane[j];
    if (!texts[t]) {
      texts[t].reset(new mapText());
      mapText* o = texts[t].get();
      o->mnId = t; o->RefKF = kfs[t_host[j]].get(); o->Nidx = (int)o->RefKF->mNcr.size(); o->STATE = TEXTGOOD;
      Mat31 n; n(0) = theta[3 * t]; n(1) = theta[3 * t + 1]; n(2) = theta[3 * t + 2];
      o->RefKF->mNcr.push_back(n);
      for (int c = 0; c < 4; ++c) { Vec2 r; r(0) = box[8 * t + 2 * c]; r(1) = box[8 * t + 2 * c + 1]; o->vTextDeteRay.push_back(r); }
      o->vRefFeature.resize(NL);
      for (int b = j; b < e; ++b) {
        tfeat.emplace_back(new TextFeature());
        TextFeature* tf = tfeat.back().get();
        tf->IdxToRaw = b - j;
        for (int k = 0; k < 8; ++k) { Mat31 r; r(0) = t_rays[16 * b + 2 * k]; r(1) = t_rays[16 * b + 2 * k + 1]; r(2) = 1.0; tf->neighbourRay.push_back(r); tf->neighbourNInten.push_back(t_iref[8 * b + k]); }
        for (int l = 0; l < NL; ++l) o->vRefFeature[l].push_back(tf);
      }
    }
    keyframe* kf = kfs[t_cam[j]].get();
    tobs.emplace_back(new TextObservation()); tobs.back()->obj = texts[t].get();
    kf->vObvText.push_back(tobs.back().get()); kf->vObvGoodTexts.push_back(true); kf->vObvGoodTextFeats.emplace_back((size_t)(e - j), true);
    texts[t]->vObvkeyframe[kf] = std::vector<int>();
    j = e;
  }
  map M;
  for (int k = 0; k < n_window; ++k) M.vKFs.push_back(kfs[k].get());
  for (auto& p : pts) if (p) M.vPts.push_back(p.get());
  for (auto& t : texts) if (t) M.vTexts.push_back(t.get());
  M.imapPts = n_points; M.imapText = n_planes; M.imapkfs = n_cams;
  // ---- the call ----
  Mat33 K; K(0, 0) = K4[0]; K(1, 1) = K4[1]; K(0, 2) = K4[2]; K(1, 2) = K4[3]; K(2, 2) = 1.0;
  double dScale = 1.0;   // every pyramid level of the test graph is the level-0 image (see tests/test_gpu_shim.py)
  int nLevels = nlevels_i; bool noText = false, rapid = false;
  optimizer opt(K, dScale, nLevels, noText, rapid);
  bool ok = true;
  if (mode == 0) opt.GlobalBA(&M);
  else if (mode == 1) opt.LocalBundleAdjustment(&M, M.vKFs, LOCAL);
  else if (mode == 2) opt.PoseOptim(*kfs[n_window - 1]);
  else if (mode == 3) opt.OptimizeLandmarker(&M);
  else if (mode == 4) { mapText* o = M.vTexts[0]; keyframe* F = kfs[t_cam[0]].get(); o->vObvkeyframe.erase(F); ok = opt.ThetaOptimMultiFs(*F, o); }
  // ---- results ----
  FILE* o = fopen(argv[2], "wb");
  std::vector<double> oc(7 * (size_t)n_cams), orho(n_points, 0.0), oth(3 * (size_t)n_planes, 0.0), ocov(9, 0.0);
  for (int k = 0; k < n_cams; ++k) {
    const keyframe* kf = kfs[k].get();
    // quaternion of mRcw (w >= 0 branch suffices for the synthetic poses) and tcw
    const double tr = kf->mRcw(0, 0) + kf->mRcw(1, 1) + kf->mRcw(2, 2), w = 0.5 * std::sqrt(tr + 1.0), s = 0.25 / w;
    oc[7 * k] = w; oc[7 * k + 1] = (kf->mRcw(2, 1) - kf->mRcw(1, 2)) * s; oc[7 * k + 2] = (kf->mRcw(0, 2) - kf->mRcw(2, 0)) * s; oc[7 * k + 3] = (kf->mRcw(1, 0) - kf->mRcw(0, 1)) * s;
    for (int c = 0; c < 3; ++c) oc[7 * k + 4 + c] = kf->mtcw(c);
  }
  for (int l = 0; l < n_points; ++l) if (pts[l]) orho[l] = pts[l]->rho;
  for (int t = 0; t < n_planes; ++t) if (texts[t]) { const Mat31& n = texts[t]->RefKF->mNcr[texts[t]->Nidx]; oth[3 * t] = n(0); oth[3 * t + 1] = n(1); oth[3 * t + 2] = n(2); }
  if (mode == 4) for (int c = 0; c < 9; ++c) ocov[c] = M.vTexts[0]->Covariance.m[c];
  int32_t counts[8] = {0, 0, 0, opt.last_summary.iterations, opt.last_summary.termination, ok ? 1 : 0, 0, 0};
  for (int k = 0; k < n_cams; ++k) {
    for (bool g : kfs[k]->vObvGoodPts) counts[0] += !g;
    for (bool g : kfs[k]->vObvGoodTexts) counts[2] += !g;
    for (auto& v : kfs[k]->vObvGoodTextFeats) for (bool g : v) counts[1] += !g;
  }
  fwrite(counts, 4, 8, o);
  fwrite(oc.data(), 8, oc.size(), o); fwrite(orho.data(), 8, orho.size(), o); fwrite(oth.data(), 8, oth.size(), o); fwrite(ocov.data(), 8, 9, o);
  fwrite(&opt.last_summary.final_cost, 8, 1, o);
  fclose(o);
  return 0;
}
