// Driver around the REFERENCE's own `nms` and `computeCovariance` (test infrastructure only).  The two function
// bodies are not in this repository: oracle/ref_build.sh extracts lines 161-340 of
// /root/reference/orb_slam2/src/cv/sp_extractor.cpp verbatim into oracle/_ref/gen/sppost_impl.inc (git-ignored) and
// compiles them here against oracle/ref_cv_stub.h (the container has neither OpenCV nor Eigen).  The C entry points
// below only marshal flat arrays in and out, the way SPExtractor::operator() calls the two functions (:502-508).
#include <queue>
#include <random>
#include <vector>

#include "ref_cv_stub.h"

using namespace cv;
using namespace std;

namespace orbslam {
#include "sppost_impl.inc"
}  // namespace orbslam

extern "C" {

// pts_sorted [n][2] (x, y) in descending score order; desc [n][256] or NULL.  Returns N; kps_xy [N][2], occ [H/8][W/8],
// desc_out [N][256] (if desc).  sel is recovered by the caller from desc rows.
int spref_nms(const float *pts_sorted, const float *desc, int n, int num_features, int border, int dist_thresh, int W, int H,
              float *kps_xy, int16_t *occ, float *desc_out, int max_out) {
  cv::Mat det(n, 2, CV_32FC1), d(n > 0 ? n : 1, 256, CV_32FC1, cv::Scalar(0));
  for (int i = 0; i < n; i++) { det.at<float>(i, 0) = pts_sorted[2 * i]; det.at<float>(i, 1) = pts_sorted[2 * i + 1]; }
  if (desc) for (int i = 0; i < n; i++) memcpy(d.data + i * d.step, desc + 256 * (size_t)i, 1024);
  std::vector<cv::KeyPoint> kps;
  cv::Mat descriptors, occ_grid;
  orbslam::nms(det, d, num_features, kps, descriptors, border, dist_thresh, W, H, occ_grid);
  const int N = (int)kps.size();
  if (N > max_out) return -1;
  for (int i = 0; i < N; i++) {
    kps_xy[2 * i] = kps[i].pt.x; kps_xy[2 * i + 1] = kps[i].pt.y;
    if (kps[i].size != 1.0f || kps[i].angle != -1 || kps[i].octave != 0) return -2;
    if (desc_out) memcpy(desc_out + 256 * (size_t)i, descriptors.data + i * descriptors.step, 1024);
  }
  for (int r = 0; r < H / 8; r++) memcpy(occ + r * (W / 8), occ_grid.data + r * occ_grid.step, (W / 8) * 2);
  return N;
}

void spref_covariance(const float *heat_inv, int h, int w, const float *kps_xy, int N, float *response, float *cov2, float *cov2_inv) {
  cv::Mat heat(h, w, CV_32FC1, const_cast<float *>(heat_inv));
  std::vector<cv::KeyPoint> kps;
  for (int i = 0; i < N; i++) kps.push_back(cv::KeyPoint(kps_xy[2 * i], kps_xy[2 * i + 1], 1.0f));
  std::vector<Eigen::Vector2f> c2, ci;
  std::vector<Eigen::Matrix2f> info;
  orbslam::computeCovariance(heat, kps, c2, ci, info);
  for (int i = 0; i < N; i++) {
    response[i] = kps[i].response;
    cov2[2 * i] = c2[i].x(); cov2[2 * i + 1] = c2[i].y();
    cov2_inv[2 * i] = ci[i].x(); cov2_inv[2 * i + 1] = ci[i].y();
  }
}
}
