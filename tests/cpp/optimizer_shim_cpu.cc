// CPU run of every overload in sp_orb_slam_b200/cpp/optimizer_dust.h against the fake backend (fake_spfe_dust.c = the
// oracle behind the C ABI entry).  usage: optimizer_shim_cpu <scene.bin> <out.txt>
// scene.bin: int32 rows, cols, n, H, W; float fx, fy, cx, cy; float Tcw[16]; float dust[rows*cols]; float heat[H*W];
//            float Xw[n*3]; uint8 null[n]; uint8 bad[n]
#include <cstdio>
#include <fstream>
#include <iostream>

#include "optimizer_dust.h"

using namespace orbslam;

struct MapPoint {
  cv::Mat X;
  bool in_view = false, bad = false;
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

static void dump(FILE *o, const char *tag, int n_inlier, const Frame &f, const std::vector<bool> &flags) {
  fprintf(o, "%s %d", tag, n_inlier);
  for (int i = 0; i < 16; i++) fprintf(o, " %.9g", f.mTcw.at<float>(i / 4, i % 4));
  fprintf(o, "\n");
  for (size_t i = 0; i < flags.size(); i++) fprintf(o, "%d", flags[i] ? 1 : 0);
  fprintf(o, "\n");
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  int32_t hdr[5];
  f.read(reinterpret_cast<char *>(hdr), sizeof hdr);
  const int rows = hdr[0], cols = hdr[1], n = hdr[2], H = hdr[3], W = hdr[4];
  float k[4], T[16];
  f.read(reinterpret_cast<char *>(k), sizeof k);
  f.read(reinterpret_cast<char *>(T), sizeof T);
  Frame base;
  base.fx = k[0]; base.fy = k[1]; base.cx = k[2]; base.cy = k[3];
  base.mTcw.create(4, 4, CV_32FC1); memcpy(base.mTcw.data, T, 64);
  base.dust_.create(rows, cols, CV_32FC1);
  f.read(reinterpret_cast<char *>(base.dust_.data), static_cast<std::streamsize>(rows) * cols * 4);
  base.heat_.create(H, W, CV_32FC1);
  f.read(reinterpret_cast<char *>(base.heat_.data), static_cast<std::streamsize>(H) * W * 4);
  std::vector<MapPoint> store(n);
  for (int i = 0; i < n; i++) { store[i].X.create(3, 1, CV_32FC1); f.read(reinterpret_cast<char *>(store[i].X.data), 12); }
  std::vector<uint8_t> isnull(n), bad(n);
  f.read(reinterpret_cast<char *>(isnull.data()), n);
  f.read(reinterpret_cast<char *>(bad.data()), n);
  if (!f) { std::cerr << "short scene file\n"; return 2; }
  Optimizer::SetBackend(reinterpret_cast<spfe_ctx *>(0x1));  // the fake backend ignores it
  FILE *o = fopen(argv[2], "w");
  // (1) the live overload: every entry gets an edge
  {
    Frame fr = base; fr.mTcw = base.mTcw.clone();
    std::vector<MapPoint *> mps;
    for (int i = 0; i < n; i++) mps.push_back(&store[i]);
    std::vector<bool> vis(n, false);
    const int r = Optimizer::PoseOptimizationDust(&fr, mps, vis);
    dump(o, "live", r, fr, vis);
    for (int i = 0; i < n; i++) fprintf(o, "%d %.9g %.9g\n", store[i].in_view ? 1 : 0, store[i].dust_proj_u, store[i].dust_proj_v);
  }
  for (int i = 0; i < n; i++) store[i].bad = bad[i] != 0;
  std::vector<MapPoint *> with_null;
  for (int i = 0; i < n; i++) with_null.push_back(isnull[i] ? nullptr : &store[i]);
  {  // (2) good map points only, no side outputs
    Frame fr = base; fr.mTcw = base.mTcw.clone();
    dump(o, "mps", Optimizer::PoseOptimizationDust(&fr, with_null), fr, {});
  }
  {  // (3) key frame
    Frame fr = base; fr.mTcw = base.mTcw.clone();
    KeyFrame kf; kf.N = n; kf.mps = with_null; kf.is_mp_visible_.assign(n, false);
    const int r = Optimizer::PoseOptimizationDust(&fr, &kf);
    dump(o, "kf", r, fr, kf.is_mp_visible_);
  }
  {  // (4) last frame; N smaller than the container: only the first N entries are read
    Frame fr = base; fr.mTcw = base.mTcw.clone();
    Frame last; last.N = n - 3; last.mvpMapPoints = with_null; last.is_mp_visible_.assign(n, false);
    const int r = Optimizer::PoseOptimizationDust(&fr, &last);
    dump(o, "last", r, fr, last.is_mp_visible_);
  }
  {  // (5) heat
    Frame fr = base; fr.mTcw = base.mTcw.clone();
    Frame last; last.N = n; last.mvpMapPoints = with_null;
    dump(o, "heat", Optimizer::PoseOptimizationHeat(&fr, &last), fr, {});
  }
  fclose(o);
  return 0;
}
