// Minimal stand-ins for the g2o / Eigen facilities that the reference's own EdgeSE3ProjectDustOnlyPose
// (orb_slam2/include/orb_slam/optimization/types_dust_tracking.h:22-65, orb_slam2/src/optimization/types_dust_tracking.cpp:36-141)
// touches, so that the class can be compiled VERBATIM from /root/reference in a container without g2o / Eigen
// (oracle/ref_build.sh).  TEST INFRASTRUCTURE ONLY.  Written for this purpose, not taken from either library:
// fixed-size row-major matrices with coefficient-wise products summed in index order (what Eigen's lazy product does
// for these sizes), SE3Quat::map = Eigen's quaternion rotation (uv = 2 q.vec x v; v + w uv + q.vec x uv) + translation.
#pragma once
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {
template <class T, int R, int C>
struct Matrix {
  T d[R * C];
  Matrix() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
  static Matrix Zero() { return Matrix(); }
  T &operator()(int i, int j) { return d[i * C + j]; }
  const T &operator()(int i, int j) const { return d[i * C + j]; }
  T &operator()(int i) { return d[i]; }
  const T &operator()(int i) const { return d[i]; }
  T &operator[](int i) { return d[i]; }
  const T &operator[](int i) const { return d[i]; }
};
template <class T, int R, int K, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K> &a, const Matrix<T, K, C> &b) {
  Matrix<T, R, C> o;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) {
      T s = a(i, 0) * b(0, j);
      for (int k = 1; k < K; k++) s = s + a(i, k) * b(k, j);
      o(i, j) = s;
    }
  return o;
}
typedef Matrix<double, 3, 1> Vector3d;
}  // namespace Eigen

namespace g2o {
typedef double number_t;
typedef Eigen::Matrix<number_t, 3, 1> Vector3;

class SE3Quat {
 public:
  double q[4] = {0, 0, 0, 1};  // x y z w
  double t[3] = {0, 0, 0};
  Vector3 map(const Vector3 &v) const {
    const double qx = q[0], qy = q[1], qz = q[2], qw = q[3];
    double uv0 = qy * v[2] - qz * v[1], uv1 = qz * v[0] - qx * v[2], uv2 = qx * v[1] - qy * v[0];
    uv0 += uv0; uv1 += uv1; uv2 += uv2;
    const double c0 = qy * uv2 - qz * uv1, c1 = qz * uv0 - qx * uv2, c2 = qx * uv1 - qy * uv0;
    Vector3 o;
    o[0] = ((v[0] + qw * uv0) + c0) + t[0];
    o[1] = ((v[1] + qw * uv1) + c1) + t[1];
    o[2] = ((v[2] + qw * uv2) + c2) + t[2];
    return o;
  }
};

struct HyperGraphVertex { virtual ~HyperGraphVertex() {} };
class VertexSE3Expmap : public HyperGraphVertex {
 public:
  const SE3Quat &estimate() const { return _estimate; }
  void setEstimate(const SE3Quat &e) { _estimate = e; }
 private:
  SE3Quat _estimate;
};

template <int D, typename E, typename VertexXi>
class BaseUnaryEdge {
 public:
  BaseUnaryEdge() : _vertices(1, nullptr) {}
  virtual ~BaseUnaryEdge() {}
  void setVertex(int i, HyperGraphVertex *v) { _vertices[i] = v; }
  void setLevel(int l) { _level = l; }
  int level() const { return _level; }
  const Eigen::Matrix<double, D, 1> &error() const { return _error; }
  const Eigen::Matrix<double, D, 6> &jacobianOplusXi() const { return _jacobianOplusXi; }
 protected:
  std::vector<HyperGraphVertex *> _vertices;
  Eigen::Matrix<double, D, 1> _error;
  Eigen::Matrix<double, D, 6> _jacobianOplusXi;
  int _level = 0;
};
}  // namespace g2o
