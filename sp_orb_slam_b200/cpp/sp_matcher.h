// Brute-force half of orbslam::SPMatcher (reference
// orb_slam2/include/orb_slam/cv/sp_matcher.h:16-19,48-49,86-93) over the C ABI.
// The guided searches of the reference class (projection / epipolar / Sim3 /
// fuse) stay as the reference's host code and consume our descriptors as-is.
#pragma once
#include <stdexcept>
#include <vector>

#include "mini_cv.h"
#include "spfe.h"

namespace orbslam {

class SPMatcher {
 public:
  explicit SPMatcher(float nnratio = 0.6) : mfNNratio(nnratio) {}

  // Device context used for the brute-force search; set once after the
  // extractor exists (Tracking::Tracking creates both, tracker.cpp:131-144).
  static void SetBackend(spfe_ctx *ctx) { backend() = ctx; }

  // cv::norm(a, b, NORM_L2) on two 1x256 CV_32F rows (sp_matcher.cpp:1636-1640).
  static float DescriptorDistance(const cv::Mat &a, const cv::Mat &b) { return spfe_l2(a.ptr<float>(), b.ptr<float>()); }

  // cv::BFMatcher(NORM_L2, crossCheck=true)::match(query) against `train`:
  // q2t[i] = matched train row or -1.  Rows must be contiguous 256 floats.
  static int MutualNN(const cv::Mat &query, const cv::Mat &train, std::vector<int> &q2t) {
    q2t.assign(query.rows, -1);
    if (query.rows == 0 || train.rows == 0) return 0;
    if (!backend()) throw std::runtime_error("SPMatcher: no backend set (call SPMatcher::SetBackend)");
    int rc = spfe_match_mutual_nn(backend(), query.ptr<float>(), query.rows, train.ptr<float>(), train.rows, q2t.data(), nullptr);
    if (rc != SPFE_OK) throw std::runtime_error(spfe_last_error(backend()));
    int n = 0;
    for (int v : q2t) n += v >= 0;
    return n;
  }

  // (KeyFrame*, Frame&) overload, sp_matcher.cpp:1642-1674: train = key-frame rows with a good map point,
  // query = every frame row; vpMatches12[query] = matched map point.  Written as a template over the
  // reference's KeyFrame / Frame / MapPoint types so that this header does not depend on them.
  template <class KeyFrameT, class FrameT, class MapPointT>
  int SearchByBruteForce(KeyFrameT *pKF1, FrameT &F2, std::vector<MapPointT *> &vpMatches12) {
    const std::vector<MapPointT *> mps1 = pKF1->GetMapPointMatches();
    vpMatches12.assign(F2.N, nullptr);
    std::vector<int> rows;
    for (size_t i = 0; i < mps1.size(); i++)
      if (mps1[i] && !mps1[i]->isBad()) rows.push_back(static_cast<int>(i));
    cv::Mat train = Gather(pKF1->mDescriptors, rows);
    std::vector<int> q2t;
    const int n = MutualNN(F2.mDescriptors, train, q2t);
    for (size_t q = 0; q < q2t.size(); q++)
      if (q2t[q] >= 0) vpMatches12[q] = mps1[rows[q2t[q]]];
    return n;  // the reference falls off the end here (UB); its caller ignores the value (tracker.cpp:378)
  }

  // (KeyFrame*, KeyFrame*) overload, sp_matcher_loop.cpp:334-376: both sides filtered by non-null map point,
  // vpMatches12[train row of KF1] = map point of the matched KF2 row; returns the number of matches.
  template <class KeyFrameT, class MapPointT>
  int SearchByBruteForce(KeyFrameT *pKF1, KeyFrameT *pKF2, std::vector<MapPointT *> &vpMatches12) {
    const std::vector<MapPointT *> mps1 = pKF1->GetMapPointMatches(), mps2 = pKF2->GetMapPointMatches();
    vpMatches12.assign(mps1.size(), nullptr);
    std::vector<int> rows1, rows2;
    for (size_t i = 0; i < mps1.size(); i++) if (mps1[i]) rows1.push_back(static_cast<int>(i));
    for (size_t i = 0; i < mps2.size(); i++) if (mps2[i]) rows2.push_back(static_cast<int>(i));
    cv::Mat train = Gather(pKF1->mDescriptors, rows1), query = Gather(pKF2->mDescriptors, rows2);
    std::vector<int> q2t;
    const int n = MutualNN(query, train, q2t);
    for (size_t q = 0; q < q2t.size(); q++)
      if (q2t[q] >= 0) vpMatches12[rows1[q2t[q]]] = mps2[rows2[q]];
    return n;
  }

  static const float TH_LOW, TH_HIGH;
  static const int HISTO_LENGTH;

 protected:
  static cv::Mat Gather(const cv::Mat &desc, const std::vector<int> &rows) {
    cv::Mat out(static_cast<int>(rows.size()), SPFE_DESC_DIM, CV_32FC1);
    for (size_t i = 0; i < rows.size(); i++) memcpy(out.ptr<float>(static_cast<int>(i)), desc.ptr<float>(rows[i]), SPFE_DESC_DIM * sizeof(float));
    return out;
  }
  static spfe_ctx *&backend() { static spfe_ctx *b = nullptr; return b; }
  float mfNNratio;
};

}  // namespace orbslam
