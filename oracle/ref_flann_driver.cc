// Driver around the REFERENCE's own SPMatcher::SearchForTriByFlann + CheckDistEpipolarLine (test infrastructure only).
// Not in this repository: oracle/ref_build.sh extracts, verbatim from /root/reference, into oracle/_ref/gen/
//   flann_tri.inc   orb_slam2/src/cv/sp_matcher.cpp:183-262   SearchForTriByFlann(KeyFrame*, KeyFrame*, cv::Mat F12, pairs&)
//   epi_check.inc   orb_slam2/src/cv/sp_matcher.cpp:441-469   CheckDistEpipolarLine
// and compiles them against the class skeletons below and oracle/ref_cv_stub.h.  cv::FlannBasedMatcher (third party,
// approximate KD-tree search) is replaced by an exact k-NN stand-in: what this pins is everything the reference does
// AROUND the search -- the 0.7 ratio test, the map-point / already-matched / epipole-distance / epipolar-line filters
// and the order in which pairs are claimed -- for the shim template SPMatcher::SearchForTriByFlann (cpp/sp_matcher.h).
#include <cstdint>
#include <cstring>
#include <vector>

#include "ref_cv_stub.h"

using namespace std;

namespace orbslam {
namespace common { bool verbose = false; }
struct NullLog { template <class T> NullLog &operator<<(const T &) { return *this; } };
#define LOG(x) NullLog()

class MapPoint {};
class KeyFrame {
 public:
  int N = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  cv::Mat Cw, Rcw, tcw;
  cv::Mat GetCameraCenter() { return Cw; }
  cv::Mat GetRotation() { return Rcw; }
  cv::Mat GetTranslation() { return tcw; }
  cv::Ptr<cv::FlannBasedMatcher> flann;
  cv::Mat mDescReamin;
  std::vector<size_t> mIndicesRemain;
  std::vector<MapPoint *> mps;
  MapPoint *GetMapPoint(size_t i) { return mps[i]; }
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvScaleFactors{1.0f};
  std::vector<Eigen::Vector2f> cov2_inv_;
};
class SPMatcher {
 public:
  int SearchForTriByFlann(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>> &vMatchedPairs);
  bool CheckDistEpipolarLine(const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const cv::Mat &F12, const KeyFrame *pKF2, const int idx);
};

#include "flann_tri.inc"
#include "epi_check.inc"

}  // namespace orbslam

using namespace orbslam;

namespace {
MapPoint g_mp;
void fill_kf(KeyFrame &kf, const float *desc, const uint8_t *has_mp, const float *kp, const float *cov2inv, int n) {
  kf.N = n;
  kf.mps.assign(n, nullptr);
  std::vector<int> remain;
  for (int i = 0; i < n; i++) {
    if (has_mp[i]) kf.mps[i] = &g_mp;
    else remain.push_back(i);
    kf.mvKeysUn.push_back(cv::KeyPoint(kp[2 * i], kp[2 * i + 1], 1.0f));
    kf.cov2_inv_.push_back(Eigen::Vector2f(cov2inv[2 * i], cov2inv[2 * i + 1]));
  }
  // KeyFrame::buildIndexes (keyframe.cpp:487-511): the rows without a map point, in index order
  kf.mDescReamin = cv::Mat((int)remain.size() > 0 ? (int)remain.size() : 1, 256, CV_32FC1, cv::Scalar(0));
  for (size_t r = 0; r < remain.size(); r++) {
    memcpy(kf.mDescReamin.data + r * kf.mDescReamin.step, desc + 256 * (size_t)remain[r], 1024);
    kf.mIndicesRemain.push_back(remain[r]);
  }
  if (remain.empty()) kf.mDescReamin = cv::Mat();
  kf.flann.p = std::make_shared<cv::FlannBasedMatcher>();
  kf.flann->add(kf.mDescReamin);
  kf.flann->train();
}
cv::Mat mat(const float *v, int r, int c) {
  cv::Mat m(r, c, CV_32FC1);
  for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.at<float>(i, j) = v[i * c + j];
  return m;
}
}  // namespace

extern "C" {
// pairs[2 * k] = row of KF1, pairs[2 * k + 1] = row of KF2 (room for N1 pairs), *npairs of them; returns nmatches
int spref_search_tri_flann(const float *desc1, const uint8_t *has_mp1, const float *kp1, const float *cov1, int n1,
                           const float *desc2, const uint8_t *has_mp2, const float *kp2, const float *cov2, int n2,
                           const float *F12, const float *Cw1, const float *R2w, const float *t2w, const float *intr2, int64_t *pairs, int *npairs) {
  KeyFrame k1, k2;
  fill_kf(k1, desc1, has_mp1, kp1, cov1, n1);
  fill_kf(k2, desc2, has_mp2, kp2, cov2, n2);
  k1.Cw = mat(Cw1, 3, 1);
  k2.Rcw = mat(R2w, 3, 3);
  k2.tcw = mat(t2w, 3, 1);
  k2.fx = intr2[0]; k2.fy = intr2[1]; k2.cx = intr2[2]; k2.cy = intr2[3];
  std::vector<std::pair<size_t, size_t>> out;
  SPMatcher m;
  const int n = m.SearchForTriByFlann(&k1, &k2, mat(F12, 3, 3), out);
  for (size_t k = 0; k < out.size(); k++) { pairs[2 * k] = (int64_t)out[k].first; pairs[2 * k + 1] = (int64_t)out[k].second; }
  *npairs = (int)out.size();  // (a KF1 row claimed twice is counted twice by nmatches but listed once, as upstream)
  return n;
}
}
