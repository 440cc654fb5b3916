// Driver around the REFERENCE's own EdgeSE3ProjectDustOnlyPose (test infrastructure only).  The class and its member
// functions are not in this repository: oracle/ref_build.sh extracts types_dust_tracking.h:22-65 and
// types_dust_tracking.cpp:36-141 verbatim from /root/reference into oracle/_ref/gen/ (git-ignored) and compiles them
// here against oracle/ref_g2o_stub.h + oracle/ref_cv_stub.h (the container has no g2o / Eigen / OpenCV).  The entry
// point runs computeError() then linearizeOplus() on one edge per map point, the way g2o's computeActiveErrors /
// buildSystem would at a fixed vertex estimate.
#include "ref_cv_stub.h"
#include "ref_g2o_stub.h"

using namespace std;

namespace g2o {
#include "dust_edge_decl.inc"
#include "dust_edge_impl.inc"
}  // namespace g2o

extern "C" int spref_dust_edges(const float *dust, int rows, int cols, const double *pose7, const double *Xw, int n, double fx, double fy,
                                double cx, double cy, unsigned char *level, double *err, float *uv, double *J) {
  cv::Mat map(rows, cols, CV_32FC1, const_cast<float *>(dust));
  g2o::VertexSE3Expmap v;
  g2o::SE3Quat T;
  for (int k = 0; k < 4; k++) T.q[k] = pose7[k];
  for (int k = 0; k < 3; k++) T.t[k] = pose7[4 + k];
  v.setEstimate(T);
  int thrown = 0;
  for (int i = 0; i < n; i++) {
    g2o::EdgeSE3ProjectDustOnlyPose e;
    e.setVertex(0, &v);
    e.fx = fx; e.fy = fy; e.cx = cx; e.cy = cy;   // optimizer_dust.cpp:222-225 (values prepared by the caller)
    e.setDustData(&map);
    e.Xw[0] = Xw[3 * i]; e.Xw[1] = Xw[3 * i + 1]; e.Xw[2] = Xw[3 * i + 2];
    e.u_ = uv[2 * i]; e.v_ = uv[2 * i + 1];        // uninitialised members upstream; only read where set
    e.setLevel(level[i]);
    e.computeError();
    err[i] = e.error()(0, 0);
    uv[2 * i] = e.u_; uv[2 * i + 1] = e.v_;
    try {
      e.linearizeOplus();
      for (int k = 0; k < 6; k++) J[6 * i + k] = e.jacobianOplusXi()(0, k);
    } catch (const std::runtime_error &) {
      for (int k = 0; k < 6; k++) J[6 * i + k] = 0.0;
      thrown = -1;
    }
    level[i] = (unsigned char)e.level();
  }
  return thrown;
}
