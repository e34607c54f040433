// Minimal stand-ins for the TextSLAM / Eigen / OpenCV types the optimizer and ORBextractor class surfaces touch, so that the
// drop-in shims (optimizer_b200.cc, ORBextractor_b200.cc) compile and run here without Ceres, Eigen, OpenCV or glog (none of
// them is installed in this image). Member names and meanings follow the reference headers:
//   src/setting.h:48-210 (SceneObservation, SceneFeature, TextObservation, TextFeature, TextStatus, BAStatus),
//   src/keyframe.h:94-155, src/frame.h, src/mapPts.h / mapPts.cc:49-69, src/mapText.h, src/map.h.
// Against the real headers the shims need no change except this include.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace cv {
struct Mat {               // CV_8UC1 only
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;
  std::vector<unsigned char> own;
  Mat() = default;
  Mat(int r, int c, unsigned char* d, size_t s = 0) : rows(r), cols(c), data(d), step(s ? s : (size_t)c) {}
  void create(int r, int c) { rows = r; cols = c; own.assign((size_t)r * c, 0); data = own.data(); step = (size_t)c; }
  bool empty() const { return data == nullptr; }
};
struct Point2f { float x = 0, y = 0; };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };   // 28 bytes like cv::KeyPoint
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace cv

namespace TextSLAM {

template <int R, int C>
struct Mat_ {
  double m[R * C];
  Mat_() { for (double& v : m) v = 0.0; }
  double& operator()(int i, int j) { return m[i * C + j]; }
  double operator()(int i, int j) const { return m[i * C + j]; }
  double& operator()(int i) { return m[i]; }
  double operator()(int i) const { return m[i]; }
  void setIdentity() { for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) m[i * C + j] = i == j ? 1.0 : 0.0; }
};
typedef Mat_<3, 3> Mat33;
typedef Mat_<3, 1> Mat31;
typedef Mat_<4, 4> Mat44;
typedef Mat_<2, 1> Vec2;
typedef Mat_<3, 1> Vec3;

enum TextStatus { TEXTGOOD = 0, TEXTIMMATURE = 1, TEXTBAD = 2 };
enum BAStatus { NOTREACHWIN = 0, LOCAL = 1 };

class keyframe;
class mapPts {
 public:
  int mnId = 0;
  keyframe* RefKF = nullptr;
  Vec3 ray;          // (x, y, 1) in the host keyframe
  double rho = 1.0;  // inverse depth
  Vec3 GetRaydir() const { return ray; }
  Vec3 GetPtInv() const { Vec3 v; v(0) = ray(0); v(1) = ray(1); v(2) = rho; return v; }   // (x, y, rho), src/mapPts.cc:49-69
  double GetInverD() const { return rho; }
  void SetRho(double r) { rho = r; }
};
struct SceneObservation { mapPts* pt = nullptr; };
struct SceneFeature { Vec2 feature; int IdxToRaw = 0; };
struct TextFeature { std::vector<Mat31> neighbourRay; std::vector<double> neighbourNInten; int IdxToRaw = 0; };
class mapText {
 public:
  int mnId = 0;
  keyframe* RefKF = nullptr;
  int Nidx = 0;
  TextStatus STATE = TEXTGOOD;
  std::vector<std::vector<TextFeature*>> vRefFeature;   // per pyramid level
  std::vector<Vec2> vTextDeteRay;                       // the four box corners as rays of the host keyframe
  std::map<keyframe*, std::vector<int>> vObvkeyframe;   // keyframes that observe the object
  Mat33 Covariance;
  int GetNidx() const { return Nidx; }
};
struct TextObservation { mapText* obj = nullptr; };

class keyframe {
 public:
  int mnId = 0;
  int iScaleLevels = 4;
  Mat33 mRcw; Mat31 mtcw; Mat44 mTcw, mTwc;
  std::vector<Mat31> mNcr;                               // plane parameters of the text objects hosted here
  std::vector<cv::Mat> vFrameImg;                        // direct-method pyramid
  std::vector<SceneObservation*> vObvPts;
  std::vector<std::vector<SceneFeature*>> vSceneObv2d;   // per pyramid level
  std::vector<TextObservation*> vObvText;
  std::vector<bool> vObvGoodPts, vObvGoodTexts;
  std::vector<std::vector<bool>> vObvGoodTextFeats;
  void SetPose(const Mat44& Tcw);
  void SetN(const Mat31& n, int idx) { mNcr[(size_t)idx] = n; }
  std::vector<TextObservation*> GetStateTextObvs(TextStatus s, std::vector<int>& vNew2Raw) const {
    std::vector<TextObservation*> out;
    vNew2Raw.clear();
    for (size_t i = 0; i < vObvText.size(); ++i) if (vObvText[i]->obj->STATE == s) { out.push_back(vObvText[i]); vNew2Raw.push_back((int)i); }
    return out;
  }
};
inline void keyframe::SetPose(const Mat44& Tcw) {
  mTcw = Tcw;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) mRcw(i, j) = Tcw(i, j); mtcw(i) = Tcw(i, 3); }
  mTwc.setIdentity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) mTwc(i, j) = mRcw(j, i);
    mTwc(i, 3) = -(mRcw(0, i) * mtcw(0) + mRcw(1, i) * mtcw(1) + mRcw(2, i) * mtcw(2));
  }
}
// the members of `frame` the optimizer reads are the keyframe's (src/frame.h)
typedef keyframe frame;

class map {
 public:
  std::vector<keyframe*> vKFs;
  std::vector<mapPts*> vPts;
  std::vector<mapText*> vTexts;
  int imapPts = 0, imapText = 0, imapkfs = 0;   // id counters (sizes of the mnId -> index tables)
  std::vector<keyframe*> GetAllKeyFrame() const { return vKFs; }
  std::vector<mapPts*> GetAllMapPoints(bool = false) const { return vPts; }
  std::vector<mapText*> GetAllMapTexts(TextStatus s) const { std::vector<mapText*> o; for (mapText* t : vTexts) if (t->STATE == s) o.push_back(t); return o; }
};

}  // namespace TextSLAM
