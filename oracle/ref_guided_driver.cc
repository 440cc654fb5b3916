// Driver around the REFERENCE's own guided-search loops (test infrastructure only).  Not in this repository:
// oracle/ref_build.sh extracts, verbatim from /root/reference, into oracle/_ref/gen/ (git-ignored)
//   frame_area.inc      orb_slam2/src/type/frame.cpp:382-474             Frame::GetFeaturesInArea
//   matcher_proj.inc    orb_slam2/src/cv/sp_matcher.cpp:344-439          SPMatcher::SearchByProjection(Frame&, MapPoints, th, th_dist), RadiusByViewingCos
//   matcher_dist.inc    orb_slam2/src/cv/sp_matcher.cpp:1636-1640        SPMatcher::DescriptorDistance
//   matcher_const.inc   orb_slam2/src/cv/sp_matcher.cpp:18-20            TH_HIGH / TH_LOW / HISTO_LENGTH
//   matcher_last.inc    orb_slam2/src/cv/sp_matcher.cpp:1439-1543        SPMatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono)
//   dust_assoc.inc      orb_slam2/src/tracking/tracker_dust.cpp:105-172  the patch-wise association block of Tracking::trackFrameDustKFLocal
// and compiles them against the class skeletons below (only the members those bodies touch; the real classes need
// ROS / g2o / OpenCV) and oracle/ref_cv_stub.h.
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "ref_cv_stub.h"

using namespace std;

namespace orbslam {

namespace common { bool verbose = false; }
namespace tracking {
bool scale_check = false;
namespace map { bool match_adaptive = false; }
namespace dust { float c2_thresh = 0.0f; int th_nmatch = 0; }
}  // namespace tracking

class MapPoint {
 public:
  bool mbTrackInView = true, bad = false, in_view = true, dust_match = false;
  int mnTrackScaleLevel = 0, nobs = 1;
  float mTrackViewCos = 1.f, mTrackProjX = 0, mTrackProjY = 0, dust_proj_u = 0, dust_proj_v = 0;
  cv::Mat desc;
  bool isBad() const { return bad; }
  int Observations() const { return nobs; }
  cv::Mat getDescTrack() const { return desc; }
  cv::Mat Xw;  // 3x1 CV_32F
  cv::Mat GetWorldPos() const { return Xw.clone(); }
};

class Frame {
 public:
  vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1) const;
  float mnMinX = 0, mnMinY = 0;
  int grid_cols = 0, grid_rows = 0, N = 0;
  cv::Mat occ_grid, mDescriptors;
  vector<cv::KeyPoint> mvKeysUn;
  vector<MapPoint *> mvpMapPoints;
  vector<float> mvScaleFactors{1.0f};
  // SearchByProjection(Cur, Last) also reads:
  cv::Mat mTcw;
  float fx = 1, fy = 1, cx = 0, cy = 0, mb = 0, mbf = 0, mnMaxX = 0, mnMaxY = 0;
  vector<bool> mvbOutlier;
  vector<cv::KeyPoint> mvKeys;
  vector<float> mvuRight;
};

class SPMatcher {
 public:
  int SearchByProjection(Frame &F, const vector<MapPoint *> &vpMapPoints, const float th, const float th_dist);
  int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);
  static const float TH_LOW, TH_HIGH;
  static const int HISTO_LENGTH;
  static float DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
  float RadiusByViewingCos(const float &viewCos);
};

#include "frame_area.inc"
#include "matcher_proj.inc"
#include "matcher_dist.inc"
#include "matcher_const.inc"
#include "matcher_last.inc"

// tracker_dust.cpp:105-172 lives inside Tracking::trackFrameDustKFLocal; the block reads `mCurrentFrame` and
// `mps_for_track` and leaves its count in `n_matches`
static int dust_association(Frame &mCurrentFrame, vector<MapPoint *> &mps_for_track) {
#include "dust_assoc.inc"
  return n_matches;
}

}  // namespace orbslam

using namespace orbslam;

namespace {
void fill_frame(Frame &F, const float *kdesc, const float *kp_un, int n, const int16_t *occ, int grid_rows, int grid_cols, float min_x, float min_y) {
  F.N = n; F.grid_rows = grid_rows; F.grid_cols = grid_cols; F.mnMinX = min_x; F.mnMinY = min_y;
  F.mDescriptors = cv::Mat(n > 0 ? n : 1, 256, CV_32FC1, cv::Scalar(0));
  for (int k = 0; k < n; k++) {
    memcpy(F.mDescriptors.data + k * F.mDescriptors.step, kdesc + 256 * (size_t)k, 1024);
    F.mvKeysUn.push_back(cv::KeyPoint(kp_un ? kp_un[2 * k] : 0.f, kp_un ? kp_un[2 * k + 1] : 0.f, 1.0f));
  }
  F.occ_grid = cv::Mat(grid_rows, grid_cols, CV_16SC1, cv::Scalar(-1));
  for (int r = 0; r < grid_rows; r++) memcpy(F.occ_grid.data + r * F.occ_grid.step, occ + r * grid_cols, grid_cols * 2);
  F.mvpMapPoints.assign(n, nullptr);
}
}  // namespace

extern "C" {

int spref_features_in_area(const int16_t *occ, int grid_rows, int grid_cols, const float *kp_un, int n, float x, float y, float r,
                           float min_x, float min_y, int32_t *out) {
  Frame F;
  std::vector<float> kd(256 * (size_t)(n > 0 ? n : 1), 0.f);
  fill_frame(F, kd.data(), kp_un, n, occ, grid_rows, grid_cols, min_x, min_y);
  const vector<size_t> v = F.GetFeaturesInArea(x, y, r);
  for (size_t i = 0; i < v.size(); i++) out[i] = (int32_t)v[i];
  return (int)v.size();
}

// SearchByProjection(Frame&, MapPoints, th, th_dist).  Per map point: desc, proj (x, y), view_cos, in_view, bad, nobs.
// kp_taken[k]: the keypoint carries an observed map point on entry.  kp2mp[k] = index of the map point assigned to
// keypoint k by the call, or -1.  Returns nmatches.
int spref_search_by_projection(int m, const float *qdesc, const float *qxy, const float *view_cos, const uint8_t *in_view, const uint8_t *bad,
                               const int32_t *nobs, const float *kdesc, const float *kp_un, int n, const int16_t *occ, int grid_rows,
                               int grid_cols, const uint8_t *kp_taken, float min_x, float min_y, float th, float th_dist,
                               float c2_adaptive, int32_t *kp2mp) {
  Frame F;
  fill_frame(F, kdesc, kp_un, n, occ, grid_rows, grid_cols, min_x, min_y);
  MapPoint holder;  // an observed map point already sitting on a keypoint
  holder.nobs = 1;
  for (int k = 0; k < n; k++) if (kp_taken && kp_taken[k]) F.mvpMapPoints[k] = &holder;
  std::vector<MapPoint> mps(m);
  std::vector<MapPoint *> vp;
  for (int i = 0; i < m; i++) {
    mps[i].desc = cv::Mat(1, 256, CV_32FC1);
    memcpy(mps[i].desc.data, qdesc + 256 * (size_t)i, 1024);
    mps[i].mTrackProjX = qxy[2 * i]; mps[i].mTrackProjY = qxy[2 * i + 1]; mps[i].mTrackViewCos = view_cos[i];
    mps[i].mbTrackInView = in_view ? in_view[i] != 0 : true; mps[i].bad = bad ? bad[i] != 0 : false; mps[i].nobs = nobs ? nobs[i] : 1;
    vp.push_back(&mps[i]);
  }
  tracking::map::match_adaptive = c2_adaptive > 0.0f;
  tracking::dust::c2_thresh = c2_adaptive;
  SPMatcher matcher;
  const int nm = matcher.SearchByProjection(F, vp, th, th_dist);
  for (int k = 0; k < n; k++) kp2mp[k] = (F.mvpMapPoints[k] && F.mvpMapPoints[k] != &holder) ? (int32_t)(F.mvpMapPoints[k] - mps.data()) : -1;
  return nm;
}

// SearchByProjection(Cur, Last, th, bMono = true).  The last frame's map points (one per last-frame keypoint, or none:
// has_mp[i] = 0) carry world positions Xw; Tcw_cur / Tcw_last are 4x4 row-major; K = fx, fy, cx, cy; bounds = mnMinX,
// mnMaxX, mnMinY, mnMaxY of the current frame.  kp2mp[k] = last-frame index whose map point landed on keypoint k.
int spref_search_by_projection_last(int m, const float *qdesc, const float *Xw, const uint8_t *has_mp, const uint8_t *outlier, const int32_t *nobs,
                                    const float *Tcw_cur, const float *Tcw_last, const float *K, const float *bounds, const float *kdesc,
                                    const float *kp_un, int n, const int16_t *occ, int grid_rows, int grid_cols, const uint8_t *kp_taken,
                                    float th, int32_t *kp2mp) {
  Frame Cur, Last;
  fill_frame(Cur, kdesc, kp_un, n, occ, grid_rows, grid_cols, bounds[0], bounds[2]);
  Cur.mnMaxX = bounds[1]; Cur.mnMaxY = bounds[3];
  Cur.fx = K[0]; Cur.fy = K[1]; Cur.cx = K[2]; Cur.cy = K[3];
  Cur.mTcw = cv::Mat(4, 4, CV_32FC1); memcpy(Cur.mTcw.data, Tcw_cur, 64);
  Last.mTcw = cv::Mat(4, 4, CV_32FC1); memcpy(Last.mTcw.data, Tcw_last, 64);
  Cur.mvuRight.assign(n, -1.0f);  // monocular
  MapPoint holder;
  holder.nobs = 1;
  for (int k = 0; k < n; k++) if (kp_taken && kp_taken[k]) Cur.mvpMapPoints[k] = &holder;
  std::vector<MapPoint> mps(m);
  Last.N = m;
  for (int i = 0; i < m; i++) {
    mps[i].desc = cv::Mat(1, 256, CV_32FC1);
    memcpy(mps[i].desc.data, qdesc + 256 * (size_t)i, 1024);
    mps[i].Xw = cv::Mat(3, 1, CV_32FC1);
    for (int k = 0; k < 3; k++) mps[i].Xw.at<float>(k, 0) = Xw[3 * i + k];
    mps[i].nobs = nobs ? nobs[i] : 1;
    Last.mvpMapPoints.push_back(has_mp && !has_mp[i] ? nullptr : &mps[i]);
    Last.mvbOutlier.push_back(outlier && outlier[i]);
    Last.mvKeys.push_back(cv::KeyPoint(0.f, 0.f, 1.0f));  // octave 0
  }
  SPMatcher matcher;
  const int nm = matcher.SearchByProjection(Cur, Last, th, true);
  for (int k = 0; k < n; k++) kp2mp[k] = (Cur.mvpMapPoints[k] && Cur.mvpMapPoints[k] != &holder) ? (int32_t)(Cur.mvpMapPoints[k] - mps.data()) : -1;
  return nm;
}

// the dust-track association block: dust_proj (u, v) in occ_grid cell units, in_view, bad.  kp2mp as above; dust_match[i] out.
int spref_dust_associate(int m, const float *qdesc, const float *quv, const uint8_t *in_view, const uint8_t *bad, const float *kdesc, int n,
                         const int16_t *occ, int grid_rows, int grid_cols, int32_t *kp2mp, uint8_t *dust_match) {
  Frame F;
  fill_frame(F, kdesc, nullptr, n, occ, grid_rows, grid_cols, 0.f, 0.f);
  std::vector<MapPoint> mps(m);
  std::vector<MapPoint *> vp;
  for (int i = 0; i < m; i++) {
    mps[i].desc = cv::Mat(1, 256, CV_32FC1);
    memcpy(mps[i].desc.data, qdesc + 256 * (size_t)i, 1024);
    mps[i].dust_proj_u = quv[2 * i]; mps[i].dust_proj_v = quv[2 * i + 1];
    mps[i].in_view = in_view ? in_view[i] != 0 : true; mps[i].bad = bad ? bad[i] != 0 : false;
    vp.push_back(&mps[i]);
  }
  const int nm = dust_association(F, vp);
  for (int k = 0; k < n; k++) kp2mp[k] = F.mvpMapPoints[k] ? (int32_t)(F.mvpMapPoints[k] - mps.data()) : -1;
  for (int i = 0; i < m; i++) dust_match[i] = mps[i].dust_match;
  return nm;
}
}
