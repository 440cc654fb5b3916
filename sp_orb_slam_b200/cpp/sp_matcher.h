// Brute-force half of orbslam::SPMatcher (reference
// orb_slam2/include/orb_slam/cv/sp_matcher.h:16-19,48-49,86-93) over the C ABI, plus the cell-grid guided searches
// whose greedy candidate loop runs on the device (spfe_search_guided): SearchByProjection(Frame&, MapPoints) and the
// dust-track patch association.  The other guided searches of the reference class (epipolar / Sim3 / fuse) stay as
// the reference's host code and consume our descriptors as-is.
#pragma once
#include <cstring>
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
  static spfe_ctx *Backend() { return backend(); }

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

  // sp_matcher.cpp:434-439
  static float RadiusByViewingCos(const float &viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }

  // SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th, th_dist), sp_matcher.cpp:344-432.
  // The per-object tests of the loop are gathered into flat arrays, the candidate search + greedy assignment runs in
  // spfe_search_guided, and the assignments are applied in map-point order.  c2_adaptive = tracking::dust::c2_thresh
  // when tracking::map::match_adaptive is set, else 0; min_x / min_y = Frame::mnMinX / mnMinY.
  template <class FrameT, class MapPointT>
  int SearchByProjection(FrameT &F, const std::vector<MapPointT *> &vpMapPoints, const float th = 1.0f, const float th_dist = 0.7f,
                         const float c2_adaptive = 0.0f, const float min_x = 0.0f, const float min_y = 0.0f) {
    const int m = static_cast<int>(vpMapPoints.size()), n = F.N;
    if (m == 0 || n == 0) return 0;
    if (!backend()) throw std::runtime_error("SPMatcher: no backend set (call SPMatcher::SetBackend)");
    std::vector<float> qdesc(static_cast<size_t>(m) * SPFE_DESC_DIM, 0.f), qxy(2 * m, 0.f), qr(m, 0.f), kp_un(2 * n);
    std::vector<uint8_t> qvalid(m, 0), qblocks(m, 0), taken(n, 0);
    const bool bFactor = th != 1.0;
    for (int i = 0; i < m; i++) {
      MapPointT *pMP = vpMapPoints[i];
      if (!pMP->mbTrackInView || pMP->isBad()) continue;
      qvalid[i] = 1;
      qblocks[i] = pMP->Observations() > 0;
      float r = RadiusByViewingCos(pMP->mTrackViewCos);
      if (bFactor) r *= th;
      qr[i] = r * F.mvScaleFactors[pMP->mnTrackScaleLevel];
      qxy[2 * i] = pMP->mTrackProjX;
      qxy[2 * i + 1] = pMP->mTrackProjY;
      const cv::Mat d = pMP->getDescTrack();
      memcpy(&qdesc[static_cast<size_t>(i) * SPFE_DESC_DIM], d.template ptr<float>(), SPFE_DESC_DIM * sizeof(float));
    }
    for (int k = 0; k < n; k++) {
      kp_un[2 * k] = F.mvKeysUn[k].pt.x;
      kp_un[2 * k + 1] = F.mvKeysUn[k].pt.y;
      taken[k] = F.mvpMapPoints[k] && F.mvpMapPoints[k]->Observations() > 0;
    }
    cv::Mat kdesc = Contiguous(F.mDescriptors), occ = Contiguous(F.occ_grid);
    spfe_guided_search g;
    memset(&g, 0, sizeof g);
    g.struct_size = sizeof g; g.mode = SPFE_GUIDED_AREA; g.m = m; g.n = n;
    g.qdesc = qdesc.data(); g.qxy = qxy.data(); g.qradius = qr.data(); g.qvalid = qvalid.data(); g.qblocks = qblocks.data();
    g.kdesc = kdesc.template ptr<float>(); g.kp_un = kp_un.data(); g.occ_grid = occ.template ptr<int16_t>();
    g.grid_rows = occ.rows; g.grid_cols = occ.cols; g.kp_taken = taken.data(); g.min_x = min_x; g.min_y = min_y;
    g.best_init = 256.0f; g.th_le = th_dist; g.th_lt = 0.7f; g.c2_adaptive = c2_adaptive;
    std::vector<int32_t> q2kp(m, -1);
    std::vector<float> qdist(m, 0.f);
    if (spfe_search_guided(backend(), &g, q2kp.data(), qdist.data(), nullptr) != SPFE_OK) throw std::runtime_error(spfe_last_error(backend()));
    int nmatches = 0;
    for (int i = 0; i < m; i++)
      if (q2kp[i] >= 0) { F.mvpMapPoints[q2kp[i]] = vpMapPoints[i]; nmatches++; }
    return nmatches;
  }

  // Patch-wise association of Tracking::trackDust, tracker_dust.cpp:112-172: map points with in_view && !isBad() look
  // at the 2 x 2 occ_grid cells at floor(dust_proj_u / v); best distance < 0.75 wins and clears the cell.
  template <class FrameT, class MapPointT>
  int DustAssociate(FrameT &F, const std::vector<MapPointT *> &mps_for_track) {
    const int m = static_cast<int>(mps_for_track.size()), n = F.N;
    if (m == 0 || n == 0) return 0;
    if (!backend()) throw std::runtime_error("SPMatcher: no backend set (call SPMatcher::SetBackend)");
    std::vector<float> qdesc(static_cast<size_t>(m) * SPFE_DESC_DIM, 0.f), qxy(2 * m, 0.f);
    std::vector<uint8_t> qvalid(m, 0);
    for (int i = 0; i < m; i++) {
      MapPointT *mp = mps_for_track[i];
      if (!mp->in_view || mp->isBad()) continue;
      qvalid[i] = 1;
      qxy[2 * i] = mp->dust_proj_u;
      qxy[2 * i + 1] = mp->dust_proj_v;
      const cv::Mat d = mp->getDescTrack();
      memcpy(&qdesc[static_cast<size_t>(i) * SPFE_DESC_DIM], d.template ptr<float>(), SPFE_DESC_DIM * sizeof(float));
    }
    cv::Mat kdesc = Contiguous(F.mDescriptors), occ = Contiguous(F.occ_grid);
    spfe_guided_search g;
    memset(&g, 0, sizeof g);
    g.struct_size = sizeof g; g.mode = SPFE_GUIDED_DUST_CELLS; g.m = m; g.n = n;
    g.qdesc = qdesc.data(); g.qxy = qxy.data(); g.qvalid = qvalid.data();
    g.kdesc = kdesc.template ptr<float>(); g.occ_grid = occ.template ptr<int16_t>(); g.grid_rows = occ.rows; g.grid_cols = occ.cols;
    g.best_init = 0.75f; g.th_le = -1e30f; g.th_lt = 0.75f;
    std::vector<int32_t> q2kp(m, -1);
    std::vector<float> qdist(m, 0.f);
    if (spfe_search_guided(backend(), &g, q2kp.data(), qdist.data(), nullptr) != SPFE_OK) throw std::runtime_error(spfe_last_error(backend()));
    int n_matches = 0;
    for (int i = 0; i < m; i++)
      if (q2kp[i] >= 0) { F.mvpMapPoints[q2kp[i]] = mps_for_track[i]; mps_for_track[i]->dust_match = true; n_matches++; }
    return n_matches;
  }

  static const float TH_LOW, TH_HIGH;
  static const int HISTO_LENGTH;

 protected:
  static cv::Mat Gather(const cv::Mat &desc, const std::vector<int> &rows) {
    cv::Mat out(static_cast<int>(rows.size()), SPFE_DESC_DIM, CV_32FC1);
    for (size_t i = 0; i < rows.size(); i++) memcpy(out.ptr<float>(static_cast<int>(i)), desc.ptr<float>(rows[i]), SPFE_DESC_DIM * sizeof(float));
    return out;
  }
  static cv::Mat Contiguous(const cv::Mat &m) { return m.step == m.cols * m.elemSize() ? m : m.clone(); }
  static spfe_ctx *&backend() { static spfe_ctx *b = nullptr; return b; }
  float mfNNratio;
};

}  // namespace orbslam
