// Brute-force half of orbslam::SPMatcher (reference
// orb_slam2/include/orb_slam/cv/sp_matcher.h:16-19,48-49,86-93) over the C ABI, plus the cell-grid guided searches
// whose greedy candidate loop runs on the device (spfe_search_guided): SearchByProjection(Frame&, MapPoints) and the
// dust-track patch association.  The other guided searches of the reference class (epipolar / Sim3 / fuse) stay as
// the reference's host code and consume our descriptors as-is.
#pragma once
#include <cstring>
#include <stdexcept>
#include <utility>
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

  // Exact 2-NN of every query row among the train rows + the reference's ratio test (0.7): knn[q] = {best, second} train
  // rows, good[q] = d_best < ratio * d_second.  Replaces `flann->knnMatch(query, matches, 2)` on the key frame's KD-tree
  // (KeyFrame::buildIndexes, keyframe.cpp:487-511) -- exact where FLANN is approximate, so the matches FLANN finds with
  // the true neighbours are reproduced and the ones it misses are found.
  static void KnnRatio(const cv::Mat &query, const cv::Mat &train, float ratio, std::vector<int32_t> &idx, std::vector<float> &dist,
                       std::vector<uint8_t> &good) {
    const int nq = query.rows;
    idx.assign(2 * static_cast<size_t>(nq), -1);
    dist.assign(2 * static_cast<size_t>(nq), 0.f);
    good.assign(nq, 0);
    if (nq == 0 || train.rows == 0) return;
    if (!backend()) throw std::runtime_error("SPMatcher: no backend set (call SPMatcher::SetBackend)");
    cv::Mat q = Contiguous(query), t = Contiguous(train);
    if (spfe_match_knn2(backend(), q.ptr<float>(), nq, t.ptr<float>(), t.rows, idx.data(), dist.data()) != SPFE_OK)
      throw std::runtime_error(spfe_last_error(backend()));
    for (int i = 0; i < nq; i++) good[i] = idx[2 * i] >= 0 && idx[2 * i + 1] >= 0 && dist[2 * i] < ratio * dist[2 * i + 1];
  }

  // sp_matcher.cpp:441-469, same arithmetic: squared distance of kp2 to the epipolar line of kp1 against
  // 3.84 / min(cov2_inv) of the key point's covariance (computeCovariance's output).
  template <class KeyFrameT>
  static bool CheckDistEpipolarLine(const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const cv::Mat &F12, const KeyFrameT *pKF2, const int idx) {
    const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);
    const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
    const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
    const auto sigma = pKF2->cov2_inv_[idx];
    const float sx = sigma.x(), sy = sigma.y();
    const float factor = 1.0f / (sx < sy ? sx : sy);
    const float num = a * kp2.pt.x + b * kp2.pt.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * factor;
  }

  // SearchForTriByFlann(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vector<pair<size_t, size_t>> &), sp_matcher.cpp:183-262
  // (LocalMapping::CreateNewMapPoints, local_mapper.cpp:620-631): the unmatched descriptors of pKF2 against pKF1's, ratio
  // test, then the reference's own filters in its own order (map point on either side, already matched, epipole
  // distance, epipolar line).  `ex`, `ey` = the epipole of pKF1's centre in pKF2 (the caller's R2w * Cw + t2w projection,
  // :186-192) so that this header needs no matrix algebra.
  template <class KeyFrameT>
  int SearchForTriByFlann(KeyFrameT *pKF1, KeyFrameT *pKF2, const cv::Mat &F12, const float ex, const float ey,
                          std::vector<std::pair<size_t, size_t>> &vMatchedPairs) {
    int nmatches = 0;
    std::vector<bool> vbMatched2(pKF2->N, false);
    std::vector<int> vMatches12(pKF1->N, -1);
    std::vector<int32_t> idx;
    std::vector<float> dist;
    std::vector<uint8_t> good;
    KnnRatio(pKF2->mDescReamin, pKF1->mDescReamin, 0.7f, idx, dist, good);
    for (size_t i = 0; i < good.size(); i++) {
      if (!good[i]) continue;
      const size_t idx1 = pKF1->mIndicesRemain[idx[2 * i]];
      if (pKF1->GetMapPoint(idx1)) continue;
      const size_t idx2 = pKF2->mIndicesRemain[i];
      if (vbMatched2[idx2] || pKF2->GetMapPoint(idx2)) continue;
      const cv::KeyPoint &kp1 = pKF1->mvKeysUn[idx1];
      const cv::KeyPoint &kp2 = pKF2->mvKeysUn[idx2];
      const float distex = ex - kp2.pt.x, distey = ey - kp2.pt.y;
      if (distex * distex + distey * distey < 100 * pKF2->mvScaleFactors[kp2.octave]) continue;
      if (CheckDistEpipolarLine(kp1, kp2, F12, pKF2, static_cast<int>(idx2))) {
        vMatches12[idx1] = static_cast<int>(idx2);
        vbMatched2[idx2] = true;
        nmatches++;
      }
    }
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (size_t i = 0, iend = vMatches12.size(); i < iend; i++)
      if (vMatches12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, static_cast<size_t>(vMatches12[i])));
    return nmatches;
  }

  // The reference's own signature: the epipole of pKF1's camera centre in pKF2 as at sp_matcher.cpp:186-192
  // (C2 = R2w * Cw + t2w: CV_32F product with OpenCV's double accumulator, then the float sum).
  template <class KeyFrameT>
  int SearchForTriByFlann(KeyFrameT *pKF1, KeyFrameT *pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t>> &vMatchedPairs) {
    const cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    float C2[3];
    for (int i = 0; i < 3; i++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += static_cast<double>(R2w.template at<float>(i, k)) * Cw.template at<float>(k, 0);
      C2[i] = static_cast<float>(s) + t2w.template at<float>(i, 0);
    }
    const float invz = 1.0f / C2[2];
    const float ex = pKF2->fx * C2[0] * invz + pKF2->cx;
    const float ey = pKF2->fy * C2[1] * invz + pKF2->cy;
    return SearchForTriByFlann(pKF1, pKF2, F12, ex, ey, vMatchedPairs);
  }

  // SearchByFlann(KeyFrame *kf_db, KeyFrame *kf_qry, vector<pair<size_t, size_t>> &), sp_matcher.cpp:264-279: upstream the body
  // stops after the ratio test (nothing is pushed, no return value).  Here the survivors are delivered as
  // (row of kf_db, row of kf_qry) in query order and counted.
  template <class KeyFrameT>
  int SearchByFlann(KeyFrameT *kf_db_ptr, KeyFrameT *kf_qry_ptr, std::vector<std::pair<size_t, size_t>> &vMatchesPairs) {
    std::vector<int32_t> idx;
    std::vector<float> dist;
    std::vector<uint8_t> good;
    KnnRatio(kf_qry_ptr->mDescReamin, kf_db_ptr->mDescReamin, 0.7f, idx, dist, good);
    vMatchesPairs.clear();
    for (size_t i = 0; i < good.size(); i++)
      if (good[i]) vMatchesPairs.push_back(std::make_pair(static_cast<size_t>(kf_db_ptr->mIndicesRemain[idx[2 * i]]), static_cast<size_t>(kf_qry_ptr->mIndicesRemain[i])));
    return static_cast<int>(vMatchesPairs.size());
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
