// Minimal stand-ins for the handful of OpenCV / Eigen facilities that the reference's own `nms` and
// `computeCovariance` (orb_slam2/src/cv/sp_extractor.cpp:161-340) (and the guided-search loops, oracle/ref_guided_driver.cc) touch, so that those functions can be compiled
// VERBATIM from /root/reference in a container without OpenCV / Eigen (oracle/ref_build.sh).  TEST INFRASTRUCTURE ONLY.
// Written for this purpose, not taken from either library; only the semantics the two functions rely on are
// provided: cv::Mat as a typed 2-D array with shared row headers, copyMakeBorder(BORDER_CONSTANT), element-wise float
// Vector2f arithmetic (IEEE single precision, no reassociation).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_16UC1 2
#define CV_16SC1 3
#define CV_32FC1 5
#define CV_8UC3 16

namespace cv {
typedef unsigned char uchar;
enum { BORDER_CONSTANT = 0 };
struct Size { int width, height; Size(int w, int h) : width(w), height(h) {} };
struct Scalar {
  double v[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
  static Scalar all(double a) { return Scalar(a, a, a, a); }
};
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct KeyPoint {
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uint8_t *data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); fill(s); }
  Mat(Size sz, int type, const Scalar &s) { create(sz.height, sz.width, type); fill(s); }
  Mat(int r, int c, int type, void *ext) : rows(r), cols(c), data(static_cast<uint8_t *>(ext)), type_(type) { step = c * elemSize(); }  // external memory (driver only)
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = c * elemSize();
    store_ = std::make_shared<std::vector<uint8_t>>(static_cast<size_t>(r) * step + 8, 0xCD);  // "uninitialised"
    data = store_->data();
  }
  size_t elemSize() const { return type_ == CV_8UC1 ? 1 : type_ == CV_8UC3 ? 3 : (type_ == CV_16UC1 || type_ == CV_16SC1) ? 2 : 4; }
  int type() const { return type_; }
  template <class T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + r * step + c * sizeof(T)); }
  template <class T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + r * step + c * sizeof(T)); }
  Mat &setTo(const Scalar &s) { fill(s); return *this; }
  bool empty_() const { return !data || rows == 0 || cols == 0; }
  Mat row(int r) const { Mat m = *this; m.rows = 1; m.data = data + r * step; return m; }  // header sharing the storage
  // sub-matrix headers and the small CV_32F algebra SearchByProjection(Cur, Last) uses on 3x3 / 3x1 blocks of mTcw
  Mat rowRange(int a, int b) const { Mat m = *this; m.rows = b - a; m.data = data + a * step; return m; }
  Mat colRange(int a, int b) const { Mat m = *this; m.cols = b - a; m.data = data + a * elemSize(); return m; }
  Mat col(int c) const { return colRange(c, c + 1); }
  template <class T> T &at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  template <class T> const T &at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  Mat t() const {
    Mat m(cols, rows, type_);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c);
    return m;
  }
  Mat operator-() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(r, c) = -at<float>(r, c);
    return m;
  }
  void push_back(const Mat &r) {  // append the rows of r (same width / type); an empty Mat adopts them
    if (!data || rows == 0) { *this = r.clone(); return; }
    Mat m(rows + r.rows, cols, type_);
    for (int i = 0; i < rows; i++) memcpy(m.data + i * m.step, data + i * step, cols * elemSize());
    for (int i = 0; i < r.rows; i++) memcpy(m.data + (rows + i) * m.step, r.data + i * r.step, cols * elemSize());
    *this = m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; r++) memcpy(m.data + r * m.step, data + r * step, cols * elemSize());
    return m;
  }
  void copyTo(Mat dst) const {  // dst is a header onto existing storage of the same shape (descriptors.row(i))
    for (int r = 0; r < rows; r++) memcpy(dst.data + r * dst.step, data + r * step, cols * elemSize());
  }
 private:
  void fill(const Scalar &s) {
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        uint8_t *p = data + r * step + c * elemSize();
        switch (type_) {
          case CV_8UC1: *p = static_cast<uint8_t>(s.v[0]); break;
          case CV_8UC3: p[0] = static_cast<uint8_t>(s.v[0]); p[1] = static_cast<uint8_t>(s.v[1]); p[2] = static_cast<uint8_t>(s.v[2]); break;
          case CV_16UC1: *reinterpret_cast<uint16_t *>(p) = static_cast<uint16_t>(s.v[0]); break;
          case CV_16SC1: *reinterpret_cast<int16_t *>(p) = static_cast<int16_t>(s.v[0]); break;
          default: *reinterpret_cast<float *>(p) = static_cast<float>(s.v[0]);
        }
      }
  }
  int type_ = CV_8UC1;
  std::shared_ptr<std::vector<uint8_t>> store_;
};
inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int /*BORDER_CONSTANT*/, const Scalar &value) {
  Mat out(src.rows + top + bottom, src.cols + left + right, src.type(), value);
  for (int r = 0; r < src.rows; r++) memcpy(out.data + (r + top) * out.step + left * src.elemSize(), src.data + r * src.step, src.cols * src.elemSize());
  dst = out;
}
inline Mat operator*(const Mat &a, const Mat &b) {  // CV_32F gemm, double accumulator like OpenCV's small-matrix path
  Mat m(a.rows, b.cols, CV_32FC1);
  for (int r = 0; r < a.rows; r++)
    for (int c = 0; c < b.cols; c++) {
      double s = 0;
      for (int k = 0; k < a.cols; k++) s += (double)a.at<float>(r, k) * b.at<float>(k, c);
      m.at<float>(r, c) = (float)s;
    }
  return m;
}
inline Mat operator+(const Mat &a, const Mat &b) {
  Mat m(a.rows, a.cols, CV_32FC1);
  for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.at<float>(r, c) + b.at<float>(r, c);
  return m;
}
enum { NORM_L2 = 4 };
// cv::norm(a, b, NORM_L2) on two CV_32F rows: sqrt of the sum of squared differences, accumulated like the oracle's
// orc_l2 (the arithmetic is OpenCV's, third-party; the pin is about the callers' control flow)
inline double norm(const Mat &a, const Mat &b, int /*NORM_L2*/) {
  float s = 0.0f;
  const float *pa = reinterpret_cast<const float *>(a.data), *pb = reinterpret_cast<const float *>(b.data);
  for (int i = 0; i < a.cols; i++) { const float t = pa[i] - pb[i]; s += t * t; }
  return (double)__builtin_sqrtf(s);
}
typedef const Mat &InputArray;
typedef Mat &OutputArray;
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
template <class T> struct Ptr {
  std::shared_ptr<T> p;
  T *operator->() const { return p.get(); }
};
// cv::BFMatcher(NORM_L2, crossCheck = true): for every query row its first-index arg-min train row, kept only if that
// train row's first-index arg-min query row is the same query (the semantics SURVEY.md 8c records; pinned against
// cv2.BFMatcher in tests/test_oracle.py::test_matcher_vs_opencv).  Stand-in written for the verbatim compile of
// SPMatcher::SearchByBruteForce -- what is checked there is the reference's row filtering and index mapping.
class BFMatcher {
 public:
  static Ptr<BFMatcher> create(int /*normType*/, bool crossCheck) { Ptr<BFMatcher> r; r.p = std::make_shared<BFMatcher>(); r.p->cross_ = crossCheck; return r; }
  void add(const Mat &train) { train_ = train; }
  void train() {}
  void match(const Mat &query, std::vector<DMatch> &matches) const {
    matches.clear();
    const int nq = query.empty_() ? 0 : query.rows, nt = train_.empty_() ? 0 : train_.rows;
    std::vector<int> t2q(nt > 0 ? nt : 1, -1);
    std::vector<float> tbest(nt > 0 ? nt : 1, 3.4e38f), qbest(nq > 0 ? nq : 1, 3.4e38f);
    std::vector<int> q2t(nq > 0 ? nq : 1, -1);
    for (int i = 0; i < nq; i++)
      for (int j = 0; j < nt; j++) {
        const float d = (float)norm(query.row(i), train_.row(j), NORM_L2);
        if (d < qbest[i]) { qbest[i] = d; q2t[i] = j; }
        if (d < tbest[j]) { tbest[j] = d; t2q[j] = i; }
      }
    for (int i = 0; i < nq; i++)
      if (q2t[i] >= 0 && (!cross_ || t2q[q2t[i]] == i)) matches.push_back(DMatch{i, q2t[i], 0, qbest[i]});
  }
 private:
  Mat train_;
  bool cross_ = false;
};
// cv::FlannBasedMatcher stand-in for the verbatim compile of SPMatcher::SearchForTriByFlann: knnMatch returns the EXACT k
// nearest train rows (first index on ties), i.e. what the KD-tree search returns whenever it finds the true neighbours.
// What this pins is the reference's control flow after the search (ratio test, map-point / epipole / epipolar filters,
// the order in which pairs are claimed); exact-vs-approximate is counted separately against cv2.FlannBasedMatcher.
class FlannBasedMatcher {
 public:
  void add(const Mat &train) { train_ = train; }
  void train() {}
  void knnMatch(const Mat &query, std::vector<std::vector<DMatch>> &matches, int k) const {
    matches.clear();
    const int nq = query.empty_() ? 0 : query.rows, nt = train_.empty_() ? 0 : train_.rows;
    for (int i = 0; i < nq; i++) {
      std::vector<DMatch> row;
      for (int r = 0; r < k && r < nt; r++) {
        int best = -1;
        float bd = 3.4e38f;
        for (int j = 0; j < nt; j++) {
          bool used = false;
          for (const DMatch &m : row) used |= m.trainIdx == j;
          if (used) continue;
          const float d = (float)norm(query.row(i), train_.row(j), NORM_L2);
          if (d < bd) { bd = d; best = j; }
        }
        row.push_back(DMatch{i, best, 0, bd});
      }
      matches.push_back(row);
    }
  }
 private:
  Mat train_;
};
}  // namespace cv
#include <cmath>
inline int cvRound(double v) { return (int)std::lrint(v); }  // OpenCV: round to nearest even (cvtsd2si / lrint)

namespace Eigen {
struct Array2f {
  float v[2];
  Array2f square() const { return Array2f{{v[0] * v[0], v[1] * v[1]}}; }
};
struct Vector2f {
  float v[2];
  Vector2f() : v{0, 0} {}
  Vector2f(float a, float b) : v{a, b} {}
  Vector2f(const Array2f &a) : v{a.v[0], a.v[1]} {}
  static Vector2f Zero() { return Vector2f(0, 0); }
  float &x() { return v[0]; }
  float &y() { return v[1]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  Array2f array() const { return Array2f{{v[0], v[1]}}; }
  Vector2f operator-(const Vector2f &o) const { return Vector2f(v[0] - o.v[0], v[1] - o.v[1]); }
  Vector2f &operator+=(const Vector2f &o) { v[0] += o.v[0]; v[1] += o.v[1]; return *this; }
  struct Comma { Vector2f *t; int i; Comma operator,(float a) { t->v[i] = a; return Comma{t, i + 1}; } };
  Comma operator<<(float a) { v[0] = a; return Comma{this, 1}; }
};
inline Vector2f operator*(float s, const Vector2f &a) { return Vector2f(s * a.v[0], s * a.v[1]); }
struct Matrix2f { float m[4]; };
}  // namespace Eigen
