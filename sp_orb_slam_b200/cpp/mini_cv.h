// Minimal stand-ins for the few OpenCV / Eigen types that appear in the
// reference's SPExtractor / SPMatcher signatures, so the shim compiles in a
// container without OpenCV / Eigen.  In the real ORB-SLAM tree define
// SPFE_WITH_OPENCV and the genuine headers are used instead.
#pragma once
#ifdef SPFE_WITH_OPENCV
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_16SC1 3
#define CV_32FC1 5

namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct KeyPoint {
  Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
  KeyPoint() = default;
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t *data = nullptr;
  Mat() = default;
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void *ext, size_t step_ = 0) : rows(r), cols(c), data(static_cast<uint8_t *>(ext)), type_(type) { step = step_ ? step_ : c * elemSize(); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = c * elemSize();
    store_ = std::make_shared<std::vector<uint8_t>>(static_cast<size_t>(r) * step);
    data = store_->data();
  }
  int type() const { return type_; }
  bool empty() const { return !data || rows == 0 || cols == 0; }
  size_t elemSize() const { return type_ == CV_8UC1 ? 1 : type_ == CV_16SC1 ? 2 : 4; }
  size_t step = 0;
  template <class T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + r * step + c * sizeof(T)); }
  template <class T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + r * step + c * sizeof(T)); }
  template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + r * step); }
  template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + r * step); }
  Mat row(int r) const { Mat m = *this; m.rows = 1; m.data = data + r * step; return m; }
  Mat clone() const { Mat m; if (!empty()) { m.create(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.data + r * m.step, data + r * step, m.step); } return m; }
  void copyTo(Mat &dst) const { dst = clone(); }
  Mat &getMat() { return *this; }
  const Mat &getMat() const { return *this; }
 private:
  int type_ = CV_8UC1;
  std::shared_ptr<std::vector<uint8_t>> store_;
};
typedef const Mat &InputArray;
typedef Mat &OutputArray;
}  // namespace cv

namespace Eigen {
struct Vector2f { float v[2] = {0, 0}; Vector2f() = default; Vector2f(float a, float b) { v[0] = a; v[1] = b; } float x() const { return v[0]; } float y() const { return v[1]; } };
}  // namespace Eigen
#endif
