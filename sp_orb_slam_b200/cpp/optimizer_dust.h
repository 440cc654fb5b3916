// Optimizer::PoseOptimizationDust(Frame*, const vector<MapPoint*>&, vector<bool>&) of the reference
// (declared orb_slam2/include/orb_slam/mapping/optimizer.h:95-96, defined orb_slam2/src/mapping/optimizer_dust.cpp:170-293)
// and its siblings on the same edge type -- PoseOptimizationDust(Frame*, mps) :524-630, (Frame*, KeyFrame*) :296-413,
// (Frame*, Frame*) :632-790, PoseOptimizationHeat(Frame*, Frame*) :415-522 -- over the C ABI: instead of building a g2o graph of EdgeSE3ProjectDustOnlyPose edges and iterating on the host, the
// map points are hoisted into a flat array and the whole 40-iteration Levenberg solve is one kernel launch
// (spfe_dust_pose_optimize).  Frame and MapPoint are template parameters so the header needs only the members the
// reference function touches: Frame::{mTcw (4x4 CV_32F), fx, fy, cx, cy, dust_, SetPose(cv::Mat)} and
// MapPoint::{GetWorldPos() (3x1 CV_32F), in_view, dust_proj_u, dust_proj_v}.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

#include "mini_cv.h"
#include "spfe.h"

namespace orbslam {

namespace Optimizer {

inline spfe_ctx *&DustBackend() { static spfe_ctx *b = nullptr; return b; }
// Device context (the extractor's); set once after the extractor exists, like SPMatcher::SetBackend.
inline void SetBackend(spfe_ctx *ctx) { DustBackend() = ctx; }

// Converter::toSE3Quat (converter.cpp:36-46): float Tcw -> double R, t -> g2o::SE3Quat(R, t) = Eigen::Quaterniond(R),
// normalizeRotation().  pose7 = (qx, qy, qz, qw, tx, ty, tz).
inline void ToSE3Quat(const cv::Mat &T, double *pose7) {
  double R[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = T.at<float>(i, j);
  double *q = pose7, t = R[0][0] + R[1][1] + R[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[2][1] - R[1][2]) * t; q[1] = (R[0][2] - R[2][0]) * t; q[2] = (R[1][0] - R[0][1]) * t;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k][j] - R[j][k]) * t; q[j] = (R[j][i] + R[i][j]) * t; q[k] = (R[k][i] + R[i][k]) * t;
  }
  if (q[3] < 0) for (int k = 0; k < 4; k++) q[k] = -q[k];
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= n;
  for (int k = 0; k < 3; k++) pose7[4 + k] = T.at<float>(k, 3);
}

// Converter::toCvMat(g2o::SE3Quat) (converter.cpp:48-51): to_homogeneous_matrix() (Eigen toRotationMatrix) -> CV_32F 4x4
inline cv::Mat ToCvMat(const double *p) {
  const double x = p[0], y = p[1], z = p[2], w = p[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x,
               tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const double R[3][3] = {{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}};
  cv::Mat T(4, 4, CV_32FC1);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T.at<float>(i, j) = static_cast<float>(R[i][j]);
    T.at<float>(i, 3) = static_cast<float>(p[4 + i]);
    T.at<float>(3, i) = 0.f;
  }
  T.at<float>(3, 3) = 1.f;
  return T;
}

// One solve: the graph every PoseOptimizationDust / PoseOptimizationHeat overload builds (one SE3 vertex = pFrame->mTcw,
// one EdgeSE3ProjectDustOnlyPose per gathered map point on `map` with intrinsics K, Huber 0.9, 40 iterations) and
// pFrame->SetPose(result).  idx[k] = position of gathered point k in the caller's container; vis / uv per gathered point.
// slot / frame >= 0: the map is still resident on the device as the dense_dust map of that frame of the slot's last
// batch (the extractor call that made pFrame), so it is read in place; otherwise `map` is uploaded.
template <class FrameT>
int SolveDust(FrameT *pFrame, const std::vector<double> &Xw, int n, const cv::Mat &map, double fx, double fy, double cx, double cy,
              double chi2_inlier, std::vector<uint8_t> &vis, std::vector<float> &uv, int slot = -1, int frame = -1) {
  if (!DustBackend()) throw std::runtime_error("Optimizer: no backend set (call Optimizer::SetBackend)");
  cv::Mat dust;
  spfe_dust_pose d;
  memset(&d, 0, sizeof d);
  d.struct_size = sizeof d; d.n = n; d.Xw = Xw.data();
  if (slot >= 0 && frame >= 0) { d.slot = slot; d.frame = frame; }
  else {
    dust = map.step == map.cols * map.elemSize() ? map : map.clone();
    d.dust = dust.template ptr<float>(); d.rows = dust.rows; d.cols = dust.cols;
  }
  d.fx = fx; d.fy = fy; d.cx = cx; d.cy = cy;
  d.huber_delta = 0.9; d.chi2_inlier = chi2_inlier; d.iterations = 40;   // optimizer_dust.cpp:219, :253, :246
  double pose[7];
  ToSE3Quat(pFrame->mTcw, pose);                                          // :186
  vis.assign(n + 1, 0);
  uv.assign(2 * static_cast<size_t>(n) + 2, 0.f);
  int32_t n_inlier = 0, n_iter = 0;
  const int rc = spfe_dust_pose_optimize(DustBackend(), &d, pose, vis.data(), uv.data(), &n_inlier, &n_iter, nullptr);
  if (rc == SPFE_ERR_STATE) throw std::runtime_error(" should be omitted");  // types_dust_tracking.cpp:114-116
  if (rc != SPFE_OK) throw std::runtime_error(spfe_last_error(DustBackend()));
  pFrame->SetPose(ToCvMat(pose));                                         // :283-287
  return n_inlier;
}

// map points with `pMP && !pMP->isBad()` of a container, in order (the gather loop every overload but the first runs)
template <class MapPointT>
void GatherGood(const std::vector<MapPointT *> &mps, int N, std::vector<double> &Xw, std::vector<int> &idx) {
  Xw.clear();
  idx.clear();
  for (int i = 0; i < N; i++) {
    MapPointT *pMP = mps[i];
    if (pMP && !pMP->isBad()) {
      const cv::Mat X = pMP->GetWorldPos();
      for (int k = 0; k < 3; k++) Xw.push_back(X.template at<float>(k, 0));
      idx.push_back(i);
    }
  }
  Xw.resize(Xw.size() + 3, 0.0);  // never hand out a null pointer
}

// dust-map intrinsics of a frame (optimizer_dust.cpp:222-225: float arithmetic, then widened to number_t)
template <class FrameT>
void DustIntrinsics(const FrameT *pFrame, double &fx, double &fy, double &cx, double &cy) {
  fx = pFrame->fx / 8.0f; fy = pFrame->fy / 8.0f;
  cx = (pFrame->cx - 3.5) / 8.0f; cy = (pFrame->cy - 3.5) / 8.0f;
}

// PoseOptimizationDust(Frame*, const vector<MapPoint*>&, vector<bool>&), optimizer_dust.cpp:170-293 -- the live one
// (tracker_dust.cpp:91).  Every entry of mps gets an edge (no null / isBad test upstream, :206-209).
template <class FrameT, class MapPointT>
int PoseOptimizationDust(FrameT *pFrame, const std::vector<MapPointT *> &mps, std::vector<bool> &is_visible, int slot = -1,
                         int frame = -1) {
  const int N = static_cast<int>(mps.size());
  std::vector<double> Xw(3 * static_cast<size_t>(N) + 3);
  for (int i = 0; i < N; i++) {
    const cv::Mat X = mps[i]->GetWorldPos();                   // :229-232
    for (int k = 0; k < 3; k++) Xw[3 * i + k] = X.template at<float>(k, 0);
  }
  double fx, fy, cx, cy;
  DustIntrinsics(pFrame, fx, fy, cx, cy);
  std::vector<uint8_t> vis;
  std::vector<float> uv;
  const int n_inlier = SolveDust(pFrame, Xw, N, pFrame->dust_, fx, fy, cx, cy, 0.9, vis, uv, slot, frame);
  for (int i = 0; i < N; i++)
    if (vis[i]) {                                               // :250-265
      is_visible[i] = true;
      mps[i]->in_view = true;
      mps[i]->dust_proj_u = uv[2 * i];
      mps[i]->dust_proj_v = uv[2 * i + 1];
    }
  return n_inlier;
}

// PoseOptimizationDust(Frame*, const vector<MapPoint*>&), optimizer_dust.cpp:524-630: good map points only, no side outputs.
template <class FrameT, class MapPointT>
int PoseOptimizationDust(FrameT *pFrame, const std::vector<MapPointT *> &mps, int slot = -1, int frame = -1) {
  std::vector<double> Xw;
  std::vector<int> idx;
  GatherGood(mps, static_cast<int>(mps.size()), Xw, idx);
  double fx, fy, cx, cy;
  DustIntrinsics(pFrame, fx, fy, cx, cy);
  std::vector<uint8_t> vis;
  std::vector<float> uv;
  return SolveDust(pFrame, Xw, static_cast<int>(idx.size()), pFrame->dust_, fx, fy, cx, cy, 0.9, vis, uv, slot, frame);
}

// PoseOptimizationDust(Frame*, KeyFrame*) (:296-413) and (Frame*, Frame*) (:632-790): the good map points of the last
// (key) frame -- GetMapPointMatches() resp. mvpMapPoints, first N entries -- and is_mp_visible_[i] = true for the inliers.
template <class FrameT, class LastT, class MapPointT>
int PoseOptimizationDustFrom(FrameT *pFrame, LastT *pLastFrame, const std::vector<MapPointT *> &last_mps, int slot = -1, int frame = -1) {
  std::vector<double> Xw;
  std::vector<int> idx;
  GatherGood(last_mps, pLastFrame->N, Xw, idx);
  double fx, fy, cx, cy;
  DustIntrinsics(pFrame, fx, fy, cx, cy);
  std::vector<uint8_t> vis;
  std::vector<float> uv;
  const int n_inlier = SolveDust(pFrame, Xw, static_cast<int>(idx.size()), pFrame->dust_, fx, fy, cx, cy, 0.9, vis, uv, slot, frame);
  for (size_t k = 0; k < idx.size(); k++)
    if (vis[k]) pLastFrame->is_mp_visible_[idx[k]] = true;     // :380-386, :754-760
  return n_inlier;
}
template <class FrameT, class KeyFrameT>
auto PoseOptimizationDust(FrameT *pFrame, KeyFrameT *pLastKF, int slot = -1, int frame = -1) -> decltype(pLastKF->GetMapPointMatches(), int()) {
  return PoseOptimizationDustFrom(pFrame, pLastKF, pLastKF->GetMapPointMatches(), slot, frame);
}
template <class FrameT>
auto PoseOptimizationDust(FrameT *pFrame, FrameT *pLastFrame, int slot = -1, int frame = -1) -> decltype(pLastFrame->mvpMapPoints, int()) {
  return PoseOptimizationDustFrom(pFrame, pLastFrame, pLastFrame->mvpMapPoints, slot, frame);
}

// PoseOptimizationHeat(Frame*, Frame*) (:415-522): the same edges on the full-resolution heat_ with the frame's pixel
// intrinsics (:473-478) and chi2 > 0.02 as the outlier test (:506); no side outputs.
template <class FrameT>
int PoseOptimizationHeat(FrameT *pFrame, FrameT *pLastFrame) {
  std::vector<double> Xw;
  std::vector<int> idx;
  GatherGood(pLastFrame->mvpMapPoints, pLastFrame->N, Xw, idx);
  std::vector<uint8_t> vis;
  std::vector<float> uv;
  return SolveDust(pFrame, Xw, static_cast<int>(idx.size()), pFrame->heat_, pFrame->fx, pFrame->fy, pFrame->cx, pFrame->cy, 0.02, vis, uv);
}

}  // namespace Optimizer

}  // namespace orbslam
