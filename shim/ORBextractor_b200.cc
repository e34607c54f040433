#include "ORBextractor_b200.h"
#include <cstdio>
#include <cstdlib>

namespace TextSLAM {

static void die(const char* what) { std::fprintf(stderr, "ORBextractor(b200): %s: %s\n", what, tslam_last_error()); std::exit(-1); }

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);   // src/ORBextractor.cc:416-432
  mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) { mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor); mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
  for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
  mvImagePyramid.resize(nlevels);
  if (tslam_ctx_create(0, &ctx)) die("tslam_ctx_create");
  if (tslam_orb_create(ctx, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, 0, &orb)) die("tslam_orb_create");
}
ORBextractor::~ORBextractor() { if (orb) tslam_orb_destroy(orb); if (ctx) tslam_ctx_destroy(ctx); }

void ORBextractor::operator()(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors) {
  if (image.empty()) return;
  const int max_kp = nfeatures + 4 * nlevels + 64;   // DistributeOctTree may exceed N per level by up to 3 (src/ORBextractor.cc:685-733)
  std::vector<tslam_keypoint> kp(max_kp);
  std::vector<uint8_t> desc((size_t)max_kp * 32);
  int32_t count = 0;
  const uint8_t* ptr = image.data;
  if (tslam_orb_extract(orb, &ptr, 1, image.cols, image.rows, (int)image.step, max_kp, kp.data(), desc.data(), &count)) die("tslam_orb_extract");
  keypoints.resize(count);
  static_assert(sizeof(cv::KeyPoint) == sizeof(tslam_keypoint), "cv::KeyPoint is 7 x 4 bytes");
  for (int i = 0; i < count; ++i) {
    cv::KeyPoint& k = keypoints[i];
    k.pt.x = kp[i].x; k.pt.y = kp[i].y; k.size = kp[i].size; k.angle = kp[i].angle; k.response = kp[i].response; k.octave = kp[i].octave; k.class_id = kp[i].class_id;
  }
  descriptors.create(count, 32);   // CV_8U
  for (int i = 0; i < count; ++i) for (int b = 0; b < 32; ++b) descriptors.data[(size_t)i * descriptors.step + b] = desc[(size_t)i * 32 + b];
}

cv::Mat ORBextractor::GetPyramidLevel(int level) {
  int w = 0, h = 0;
  if (tslam_orb_level_size(orb, level, &w, &h)) die("tslam_orb_level_size");
  cv::Mat m;
  m.create(h, w);
  if (tslam_orb_get_level(orb, 0, level, m.data)) die("tslam_orb_get_level");
  return m;
}

}  // namespace TextSLAM
