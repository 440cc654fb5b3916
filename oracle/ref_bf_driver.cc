// Driver around the REFERENCE's own SPMatcher::SearchByBruteForce overloads (test infrastructure only).  Not in this
// repository: oracle/ref_build.sh extracts, verbatim from /root/reference, into oracle/_ref/gen/ (git-ignored)
//   bf_kf_frame.inc   orb_slam2/src/cv/sp_matcher.cpp:1642-1674        SearchByBruteForce(KeyFrame*, Frame&, vector<MapPoint*>&)
//   bf_kf_kf.inc      orb_slam2/src/cv/sp_matcher_loop.cpp:334-376     SearchByBruteForce(KeyFrame*, KeyFrame*, vector<MapPoint*>&)
// and compiles them against the class skeletons below and oracle/ref_cv_stub.h (whose cv::BFMatcher is a stand-in:
// what this pins is the reference's row filtering and index mapping around the matcher).  The first overload is
// declared `int` but has no return statement upstream (undefined behaviour; its caller ignores the value,
// tracker.cpp:378) and g++ plants a trap at its end, so its text is included with `int` defined as `void` -- the only
// `int` in those 33 lines is the return type.
#include <cstdint>
#include <vector>

#include "ref_cv_stub.h"

using namespace std;

namespace orbslam {

class MapPoint {
 public:
  bool bad = false;
  bool isBad() const { return bad; }
};
class KeyFrame {
 public:
  int N = 0;
  cv::Mat mDescriptors;
  vector<cv::KeyPoint> mvKeysUn;
  vector<MapPoint *> mps;
  vector<MapPoint *> GetMapPointMatches() { return mps; }
};
class Frame {
 public:
  int N = 0;
  cv::Mat mDescriptors;
  vector<cv::KeyPoint> mvKeysUn;
};
class SPMatcher {
 public:
  void SearchByBruteForce(KeyFrame *pKF1, Frame &pKF2, std::vector<MapPoint *> &vpMatches12);  // `int` upstream, see above
  int SearchByBruteForce(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12);
};

#define int void
#include "bf_kf_frame.inc"
#undef int
#include "bf_kf_kf.inc"

}  // namespace orbslam

using namespace orbslam;

namespace {
void fill_kf(KeyFrame &kf, std::vector<MapPoint> &store, const float *desc, const uint8_t *has_mp, const uint8_t *bad, int n) {
  kf.N = n;
  kf.mDescriptors = cv::Mat(n > 0 ? n : 1, 256, CV_32FC1, cv::Scalar(0));
  store.resize(n);
  for (int i = 0; i < n; i++) {
    memcpy(kf.mDescriptors.data + i * kf.mDescriptors.step, desc + 256 * (size_t)i, 1024);
    store[i].bad = bad && bad[i];
    kf.mps.push_back(has_mp && !has_mp[i] ? nullptr : &store[i]);
  }
}
}  // namespace

extern "C" {

// (KeyFrame*, Frame&): matches12[q] = key-frame row whose map point was assigned to frame row q, or -1
void spref_bruteforce_kf_frame(const float *desc1, const uint8_t *has_mp1, const uint8_t *bad1, int n1, const float *desc2, int n2, int32_t *matches12) {
  KeyFrame kf;
  std::vector<MapPoint> s1;
  fill_kf(kf, s1, desc1, has_mp1, bad1, n1);
  Frame fr;
  fr.N = n2;
  fr.mDescriptors = cv::Mat(n2 > 0 ? n2 : 1, 256, CV_32FC1, cv::Scalar(0));
  for (int i = 0; i < n2; i++) memcpy(fr.mDescriptors.data + i * fr.mDescriptors.step, desc2 + 256 * (size_t)i, 1024);
  if (n2 == 0) fr.mDescriptors = cv::Mat();
  std::vector<MapPoint *> m12;
  SPMatcher matcher;
  matcher.SearchByBruteForce(&kf, fr, m12);
  for (int q = 0; q < n2; q++) matches12[q] = m12[q] ? (int32_t)(m12[q] - s1.data()) : -1;
}

// (KeyFrame*, KeyFrame*): matches12[i] = row of KF2 whose map point was assigned to row i of KF1, or -1; returns the count
int spref_bruteforce_kf_kf(const float *desc1, const uint8_t *has_mp1, const uint8_t *bad1, int n1, const float *desc2, const uint8_t *has_mp2,
                           const uint8_t *bad2, int n2, int32_t *matches12) {
  KeyFrame k1, k2;
  std::vector<MapPoint> s1, s2;
  fill_kf(k1, s1, desc1, has_mp1, bad1, n1);
  fill_kf(k2, s2, desc2, has_mp2, bad2, n2);
  std::vector<MapPoint *> m12;
  SPMatcher matcher;
  const int n = matcher.SearchByBruteForce(&k1, &k2, m12);
  for (int i = 0; i < n1; i++) matches12[i] = m12[i] ? (int32_t)(m12[i] - s2.data()) : -1;
  return n;
}
}
