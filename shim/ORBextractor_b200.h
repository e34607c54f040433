// Drop-in replacement of TextSLAM::ORBextractor (src/ORBextractor.h:45-114): same constructor and operator(); the body of
// operator() (src/ORBextractor.cc:1054-1116: pyramid, FAST per cell, quad-tree distribution, orientation, blur, rBRIEF) is one
// call into libtslam_b200.so. mvImagePyramid is filled on demand by GetPyramidLevel (nothing outside the class reads it).
#pragma once
#ifdef TSLAM_SHIM_STUB_TYPES
#include "stub/textslam_stub.h"
#else
#include <opencv2/core.hpp>
#endif
#include <vector>
#include "../include/tslam_b200.h"

namespace TextSLAM {

class ORBextractor {
 public:
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();
  // mask is ignored, like in the reference (src/ORBextractor.cc:1054-1061); image must be CV_8UC1
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);
  int GetLevels() const { return nlevels; }
  float GetScaleFactor() const { return (float)scaleFactor; }
  std::vector<float> GetScaleFactors() const { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() const { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() const { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() const { return mvInvLevelSigma2; }
  cv::Mat GetPyramidLevel(int level);   // level of the last image, without the 19-px border
  std::vector<cv::Mat> mvImagePyramid;

 protected:
  int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  tslam_ctx* ctx = nullptr;
  tslam_orb* orb = nullptr;
};

}  // namespace TextSLAM
