// Drives Optimizer::PoseOptimizationDust of the shim (optimizer_dust.h) the way Tracking::trackFrameDustKFLocal does
// (tracker_dust.cpp:91): a Frame with a pose, intrinsics and a dust map, a vector of map points.
// usage: dust_pose_selftest <weights> <scene.bin> <out.txt> [use_device_map: raw u8 frame file]
// scene.bin: int32 rows, cols, n; float fx, fy, cx, cy; float Tcw[16]; float dust[rows*cols]; float Xw[n*3]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "optimizer_dust.h"
#include "sp_extractor.h"
#include "sp_matcher.h"

using namespace orbslam;

struct MapPoint {
  cv::Mat X;
  bool in_view = false, dust_match = false, bad = false;
  float dust_proj_u = -1, dust_proj_v = -1;
  cv::Mat GetWorldPos() const { return X.clone(); }
  bool isBad() const { return bad; }
};
struct Frame {
  cv::Mat mTcw, dust_, heat_;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  int N = 0;
  std::vector<MapPoint *> mvpMapPoints;
  std::vector<bool> is_mp_visible_;
  void SetPose(cv::Mat T) { mTcw = T.clone(); }
};
struct KeyFrame {
  int N = 0;
  std::vector<MapPoint *> mps;
  std::vector<bool> is_mp_visible_;
  std::vector<MapPoint *> GetMapPointMatches() { return mps; }
};

int main(int argc, char **argv) {
  if (argc < 4) { std::cerr << "usage\n"; return 2; }
  std::ifstream f(argv[2], std::ios::binary);
  int32_t hdr[3];
  f.read(reinterpret_cast<char *>(hdr), sizeof hdr);
  const int rows = hdr[0], cols = hdr[1], n = hdr[2];
  Frame fr;
  float k[4];
  f.read(reinterpret_cast<char *>(k), sizeof k);
  fr.fx = k[0]; fr.fy = k[1]; fr.cx = k[2]; fr.cy = k[3];
  fr.mTcw.create(4, 4, CV_32FC1);
  f.read(reinterpret_cast<char *>(fr.mTcw.data), 64);
  fr.dust_.create(rows, cols, CV_32FC1);
  f.read(reinterpret_cast<char *>(fr.dust_.data), static_cast<std::streamsize>(rows) * cols * 4);
  std::vector<MapPoint> store(n);
  std::vector<MapPoint *> mps;
  for (int i = 0; i < n; i++) {
    store[i].X.create(3, 1, CV_32FC1);
    f.read(reinterpret_cast<char *>(store[i].X.data), 12);
    mps.push_back(&store[i]);
  }
  if (!f) { std::cerr << "short scene file\n"; return 2; }
  common::model_path = argv[1];
  camera::height = rows * 8;
  camera::width = cols * 8;
  SPExtractor ex(800);  // owns the device context; its constructor hands it to SPMatcher and Optimizer
  double start[7];
  Optimizer::ToSE3Quat(fr.mTcw, start);
  std::vector<bool> is_visible(n, false);
  const int n_inlier = Optimizer::PoseOptimizationDust(&fr, mps, is_visible);
  if (argc > 99) {  // the sibling overloads (optimizer_dust.cpp:296-790) are instantiated, not run, by this self-test
    KeyFrame kf;
    Frame last;
    Optimizer::PoseOptimizationDust(&fr, mps);
    Optimizer::PoseOptimizationDust(&fr, &kf);
    Optimizer::PoseOptimizationDust(&fr, &last);
    Optimizer::PoseOptimizationHeat(&fr, &last);
  }
  FILE *o = fopen(argv[3], "w");
  fprintf(o, "%d\n", n_inlier);
  for (int i = 0; i < 7; i++) fprintf(o, "%.17g ", start[i]);
  fprintf(o, "\n");
  for (int i = 0; i < 16; i++) fprintf(o, "%.9g ", fr.mTcw.at<float>(i / 4, i % 4));
  fprintf(o, "\n");
  for (int i = 0; i < n; i++) fprintf(o, "%d %d %.9g %.9g\n", is_visible[i] ? 1 : 0, store[i].in_view ? 1 : 0, store[i].dust_proj_u, store[i].dust_proj_v);
  fclose(o);
  printf("PoseOptimizationDust: %d of %d map points visible\n", n_inlier, n);
  return 0;
}
