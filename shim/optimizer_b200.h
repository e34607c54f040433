// Drop-in replacement of the solve entry points of TextSLAM::optimizer (src/optimizer.h:52-70): same class name, constructor and
// method signatures, same in-place effects on the object graph (SetPose / SetRho / SetN, the vObvGood* flags); the ceres::Problem
// construction + ceres::Solve + Problem::Evaluate + outlier loops of every Pyr* function are one call into libtslam_b200.so.
// InitBA, OptimizeSim3 and OptimizeLoop (two-view initialisation, loop-closure pose graph: tiny problems) and the
// text-label bookkeeping (UpdateTrackedText*, ShowBAReproj_TextBox) stay with the reference's own code.
#pragma once
#ifdef TSLAM_SHIM_STUB_TYPES
#include "stub/textslam_stub.h"
#else
#include <keyframe.h>
#include <map.h>
#include <frame.h>
#endif
#include "../include/tslam_b200.h"

namespace TextSLAM {

class optimizer {
 public:
  optimizer(Mat33& mK, double& dScale, int& nLevels, bool& Flag_noText, bool& Flag_rapid);
  ~optimizer();
  void PoseOptim(frame& F);
  void LocalBundleAdjustment(map* mpMap, std::vector<keyframe*> vKFs, const BAStatus& STATE);
  void GlobalBA(map* mpMap);
  void OptimizeLandmarker(map* mpMap);
  bool ThetaOptimMultiFs(const frame& F, mapText*& obj);

  tslam_solve_summary last_summary;   // Solver::Summary of the last ceres::Solve replacement (the reference discards it)
  int text_jac_mode = TSLAM_JAC_ANALYTIC;   // TSLAM_JAC_CENTRAL_DIFF = what Ceres computes for the nume_* functors

 private:
  struct Flat;
  void PyrPoseOptim(frame& F, double* pose, int PyBegin, double chi2Mono, double chi2Text, int its, std::vector<bool>& vPtsGood,
                    std::vector<bool>& vTextsGood, std::vector<std::vector<bool>>& vTextFeatsGood, const std::vector<TextObservation*>& TextObjs);
  void PyrBA(double** pose, double** theta, double** rho, const std::vector<keyframe*>& vKFs, const std::vector<int>& vmnId2Pts, const std::vector<bool>& vPtOptim,
             const std::vector<int>& vmnId2Texts, const std::vector<bool>& vTextOptim, const std::vector<int>& vmnId2vKFs, const std::vector<int>& InitialIdx,
             int PyBegin, double chi2Mono, double chi2Text, int its, const BAStatus& STATE);
  void PyrGlobalBA(double** pose, double** theta, double** rho, const std::vector<keyframe*>& vKFs, const std::vector<int>& vmnId2Pts,
                   const std::vector<int>& vmnId2Texts, const std::vector<int>& vmnId2vKFs, const std::vector<int>& InitialIdx, int PyBegin, int its);
  void PyrLandmarkers(double** pose, double** theta, double** rho, const std::vector<keyframe*>& vKFs, const std::vector<int>& vmnId2Pts,
                      const std::vector<int>& vmnId2Texts, std::vector<std::vector<bool>>& vPtsGoodkf, int PyBegin, double chi2Mono, int its);
  bool PyrThetaOptim(const std::vector<cv::Mat>& vImg, const std::vector<Mat44>& vTcr, mapText* obj, int PyBegin, double* theta, Mat33& thetaVariance);

  Mat33 K;
  std::vector<Mat33> vK;
  bool bFlag_noText, bFlag_rapid;
  tslam_ctx* ctx = nullptr;
};

}  // namespace TextSLAM
