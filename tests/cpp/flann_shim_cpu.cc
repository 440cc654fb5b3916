// CPU run of the SearchForTriByFlann / SearchByFlann templates of sp_orb_slam_b200/cpp/sp_matcher.h against the fake
// backend (fake_spfe_guided.c: spfe_match_knn2 answered by the oracle).  The pytest compares the pairs with the
// REFERENCE's own SearchForTriByFlann compiled verbatim (oracle/_ref/libspflann_ref.so).
// usage: flann_shim_cpu <scene.bin> <out.txt>
// scene.bin: int32 n1, n2; float F12[9], Cw[3], R2w[9], t2w[3], intr2[4]; then for KF1 and KF2:
//            float desc[n*256], kp[n*2], cov2inv[n*2]; uint8 has_mp[n]
#include <cstdio>
#include <fstream>
#include <iostream>

#include "sp_matcher.h"

using namespace orbslam;

struct MapPoint {};
struct KeyFrame {
  int N = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  cv::Mat Cw, Rcw, tcw, mDescReamin;
  cv::Mat GetCameraCenter() { return Cw; }
  cv::Mat GetRotation() { return Rcw; }
  cv::Mat GetTranslation() { return tcw; }
  std::vector<size_t> mIndicesRemain;
  std::vector<MapPoint *> mps;
  MapPoint *GetMapPoint(size_t i) { return mps[i]; }
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<float> mvScaleFactors{1.0f};
  std::vector<Eigen::Vector2f> cov2_inv_;
};

template <class T> static std::vector<T> rd(std::ifstream &f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char *>(v.data()), static_cast<std::streamsize>(n * sizeof(T)));
  return v;
}
static cv::Mat mat(const std::vector<float> &v, int r, int c) {
  cv::Mat m(r, c, CV_32FC1);
  for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.at<float>(i, j) = v[i * c + j];
  return m;
}
static MapPoint g_mp;
static void fill(KeyFrame &kf, std::ifstream &f, int n) {
  auto desc = rd<float>(f, 256 * (size_t)n), kp = rd<float>(f, 2 * (size_t)n), cov = rd<float>(f, 2 * (size_t)n);
  auto has = rd<uint8_t>(f, n);
  kf.N = n;
  std::vector<int> remain;
  for (int i = 0; i < n; i++) {
    kf.mps.push_back(has[i] ? &g_mp : nullptr);
    if (!has[i]) remain.push_back(i);
    kf.mvKeysUn.push_back(cv::KeyPoint(kp[2 * i], kp[2 * i + 1], 1.0f));
    kf.cov2_inv_.push_back(Eigen::Vector2f(cov[2 * i], cov[2 * i + 1]));
  }
  kf.mDescReamin.create((int)remain.size(), 256, CV_32FC1);   // KeyFrame::buildIndexes, keyframe.cpp:487-511
  for (size_t r = 0; r < remain.size(); r++) {
    memcpy(kf.mDescReamin.ptr<float>((int)r), &desc[256 * (size_t)remain[r]], 1024);
    kf.mIndicesRemain.push_back(remain[r]);
  }
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  auto hdr = rd<int32_t>(f, 2);
  auto F = rd<float>(f, 9), Cw = rd<float>(f, 3), R = rd<float>(f, 9), t = rd<float>(f, 3), intr = rd<float>(f, 4);
  KeyFrame k1, k2;
  fill(k1, f, hdr[0]);
  fill(k2, f, hdr[1]);
  if (!f) { std::cerr << "short scene file\n"; return 2; }
  k1.Cw = mat(Cw, 3, 1); k2.Rcw = mat(R, 3, 3); k2.tcw = mat(t, 3, 1);
  k2.fx = intr[0]; k2.fy = intr[1]; k2.cx = intr[2]; k2.cy = intr[3];
  SPMatcher::SetBackend(reinterpret_cast<spfe_ctx *>(0x1));
  SPMatcher matcher(0.7f);
  std::vector<std::pair<size_t, size_t>> pairs, all;
  const int n = matcher.SearchForTriByFlann(&k1, &k2, mat(F, 3, 3), pairs);
  const int na = matcher.SearchByFlann(&k1, &k2, all);
  FILE *o = fopen(argv[2], "w");
  fprintf(o, "%d\n", n);
  for (auto &p : pairs) fprintf(o, "%zu %zu ", p.first, p.second);
  fprintf(o, "\n%d\n", na);
  for (auto &p : all) fprintf(o, "%zu %zu ", p.first, p.second);
  fprintf(o, "\n");
  fclose(o);
  return 0;
}
