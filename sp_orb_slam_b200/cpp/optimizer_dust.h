// Optimizer::PoseOptimizationDust(Frame*, const vector<MapPoint*>&, vector<bool>&) of the reference
// (declared orb_slam2/include/orb_slam/mapping/optimizer.h:95-96, defined orb_slam2/src/mapping/optimizer_dust.cpp:170-293)
// over the C ABI: instead of building a g2o graph of EdgeSE3ProjectDustOnlyPose edges and iterating on the host, the
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

// slot / frame >= 0: pFrame->dust_ is still resident on the device as the dense_dust map of that frame of the slot's
// last batch (the extractor call that made pFrame), so it is read in place; otherwise pFrame->dust_ is uploaded.
template <class FrameT, class MapPointT>
int PoseOptimizationDust(FrameT *pFrame, const std::vector<MapPointT *> &mps, std::vector<bool> &is_visible, int slot = -1,
                         int frame = -1) {
  if (!DustBackend()) throw std::runtime_error("Optimizer: no backend set (call Optimizer::SetBackend)");
  const int N = static_cast<int>(mps.size());
  std::vector<double> Xw(3 * static_cast<size_t>(N) + 3);
  for (int i = 0; i < N; i++) {
    const cv::Mat X = mps[i]->GetWorldPos();                   // optimizer_dust.cpp:229-232
    for (int k = 0; k < 3; k++) Xw[3 * i + k] = X.template at<float>(k, 0);
  }
  cv::Mat dust;
  spfe_dust_pose d;
  memset(&d, 0, sizeof d);
  d.struct_size = sizeof d; d.n = N; d.Xw = Xw.data();
  if (slot >= 0 && frame >= 0) { d.slot = slot; d.frame = frame; }
  else {
    dust = pFrame->dust_.step == pFrame->dust_.cols * pFrame->dust_.elemSize() ? pFrame->dust_ : pFrame->dust_.clone();
    d.dust = dust.template ptr<float>(); d.rows = dust.rows; d.cols = dust.cols;
  }
  d.fx = pFrame->fx / 8.0f; d.fy = pFrame->fy / 8.0f;          // :222-225 (float arithmetic, then widened to number_t)
  d.cx = (pFrame->cx - 3.5) / 8.0f; d.cy = (pFrame->cy - 3.5) / 8.0f;
  d.huber_delta = 0.9; d.chi2_inlier = 0.9; d.iterations = 40; // :219, :253, :246
  double pose[7];
  ToSE3Quat(pFrame->mTcw, pose);                                // :186
  std::vector<uint8_t> vis(N + 1, 0);
  std::vector<float> uv(2 * static_cast<size_t>(N) + 2, 0.f);
  int32_t n_inlier = 0, n_iter = 0;
  const int rc = spfe_dust_pose_optimize(DustBackend(), &d, pose, vis.data(), uv.data(), &n_inlier, &n_iter, nullptr);
  if (rc == SPFE_ERR_STATE) throw std::runtime_error(" should be omitted");  // types_dust_tracking.cpp:114-116
  if (rc != SPFE_OK) throw std::runtime_error(spfe_last_error(DustBackend()));
  for (int i = 0; i < N; i++)
    if (vis[i]) {                                               // :250-265
      is_visible[i] = true;
      mps[i]->in_view = true;
      mps[i]->dust_proj_u = uv[2 * i];
      mps[i]->dust_proj_v = uv[2 * i + 1];
    }
  pFrame->SetPose(ToCvMat(pose));                               // :283-287
  return n_inlier;
}

}  // namespace Optimizer

}  // namespace orbslam
