// Exercises the drop-in classes the way the reference's callers do
// (Frame::ExtractORB, frame.cpp:296-314; trackReferenceKeyFrameANN, tracker.cpp:372-417).
// usage: shim_selftest <weights> <H> <W> <raw u8 frame A> <raw u8 frame B> <out prefix>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "sp_extractor.h"
#include "sp_matcher.h"

using namespace orbslam;

struct MapPoint {
  bool bad = false; bool isBad() const { return bad; }
  // the fields SearchByProjection(Frame&, MapPoints) / the dust association read (map_point.h)
  bool mbTrackInView = true, in_view = true, dust_match = false;
  float mTrackProjX = 0, mTrackProjY = 0, mTrackViewCos = 1.f, dust_proj_u = 0, dust_proj_v = 0;
  int mnTrackScaleLevel = 0, nobs = 1;
  cv::Mat desc;
  int Observations() const { return nobs; }
  cv::Mat getDescTrack() const { return desc; }
};
struct KeyFrame {
  cv::Mat mDescriptors; std::vector<MapPoint *> mps;
  std::vector<MapPoint *> GetMapPointMatches() { return mps; }
};
struct Frame {
  cv::Mat mDescriptors, occ_grid; int N = 0;
  std::vector<cv::KeyPoint> mvKeysUn; std::vector<MapPoint *> mvpMapPoints; std::vector<float> mvScaleFactors{1.0f};
};

static cv::Mat read_raw(const char *path, int H, int W) {
  cv::Mat m(H, W, CV_8UC1);
  std::ifstream f(path, std::ios::binary);
  f.read(reinterpret_cast<char *>(m.data), static_cast<std::streamsize>(H) * W);
  if (!f) { std::cerr << "cannot read " << path << "\n"; exit(2); }
  return m;
}

int main(int argc, char **argv) {
  if (argc < 7) { std::cerr << "usage\n"; return 2; }
  common::model_path = argv[1];
  camera::height = atoi(argv[2]);
  camera::width = atoi(argv[3]);
  cv::Mat a = read_raw(argv[4], camera::height, camera::width), b = read_raw(argv[5], camera::height, camera::width);
  BaseExtractor *ex = new SPExtractor(800);
  std::vector<cv::KeyPoint> ka, kb;
  cv::Mat da, db;
  (*ex)(a, cv::Mat(), ka, da);
  // what Frame::ExtractORB reads after the call
  SPExtractor *sp = dynamic_cast<SPExtractor *>(ex);
  auto cov = sp->getCov2Inv();
  cv::Mat dust = sp->dense_dust_.clone(), occ = sp->occ_grid_.clone(), heat = sp->heat_.clone();
  (*ex)(b, cv::Mat(), kb, db);
  printf("levels %d scale %.1f  A: %zu kps  B: %zu kps  cov %zu  dust %dx%d occ %dx%d heat %dx%d\n", ex->GetLevels(), ex->GetScaleFactor(),
         ka.size(), kb.size(), cov.size(), dust.rows, dust.cols, occ.rows, occ.cols, heat.rows, heat.cols);
  // key frame = frame A where every third keypoint has no map point and every tenth is bad
  KeyFrame kf; kf.mDescriptors = da;
  std::vector<MapPoint> store(ka.size());
  for (size_t i = 0; i < ka.size(); i++) { store[i].bad = (i % 10 == 9); kf.mps.push_back(i % 3 == 2 ? nullptr : &store[i]); }
  Frame fr; fr.mDescriptors = db; fr.N = static_cast<int>(kb.size());
  SPMatcher matcher(0.7f);
  std::vector<MapPoint *> m12;
  int n = matcher.SearchByBruteForce(&kf, fr, m12);
  KeyFrame kf2; kf2.mDescriptors = db; std::vector<MapPoint> store2(kb.size());
  for (size_t i = 0; i < kb.size(); i++) kf2.mps.push_back(i % 4 == 3 ? nullptr : &store2[i]);
  std::vector<MapPoint *> mkk;
  int n2 = matcher.SearchByBruteForce(&kf, &kf2, mkk);
  printf("SearchByBruteForce(KF,Frame) %d matches; (KF,KF) %d matches; d(0,0)=%.6f\n", n, n2,
         ka.empty() || kb.empty() ? 0.f : SPMatcher::DescriptorDistance(da.row(0), db.row(0)));
  // guided searches: map points = frame A's keypoints projected where they were seen, searched in frame B
  std::vector<MapPoint> mp3(ka.size());
  std::vector<MapPoint *> vmp;
  for (size_t i = 0; i < ka.size(); i++) {
    mp3[i].desc = da.row(static_cast<int>(i)).clone();
    mp3[i].mTrackProjX = ka[i].pt.x; mp3[i].mTrackProjY = ka[i].pt.y;
    mp3[i].dust_proj_u = (ka[i].pt.x - 3.5f) / 8.0f; mp3[i].dust_proj_v = (ka[i].pt.y - 3.5f) / 8.0f;
    mp3[i].mTrackViewCos = (i % 2) ? 0.9995f : 0.99f;
    mp3[i].mbTrackInView = mp3[i].in_view = (i % 7 != 6);
    mp3[i].nobs = (i % 5 == 4) ? 0 : 2;
    vmp.push_back(&mp3[i]);
  }
  (*ex)(b, cv::Mat(), kb, db);  // occ_grid_ of frame B
  fr.occ_grid = sp->occ_grid_.clone(); fr.mvKeysUn = kb; fr.mvpMapPoints.assign(kb.size(), nullptr);
  const int n3 = matcher.SearchByProjection(fr, vmp, 3.0f, 0.7f);
  std::vector<MapPoint *> proj_assign = fr.mvpMapPoints;
  fr.mvpMapPoints.assign(kb.size(), nullptr);
  const int n4 = matcher.DustAssociate(fr, vmp);
  printf("SearchByProjection(F,MPs) %d matches; dust association %d matches\n", n3, n4);
  // dump for the python-side check
  std::string pre = argv[6];
  {
    FILE *g = fopen((pre + "_guided.txt").c_str(), "w");
    for (size_t k = 0; k < kb.size(); k++)
      fprintf(g, "%ld %ld\n", proj_assign[k] ? static_cast<long>(proj_assign[k] - mp3.data()) : -1L,
              fr.mvpMapPoints[k] ? static_cast<long>(fr.mvpMapPoints[k] - mp3.data()) : -1L);
    fclose(g);
  }
  FILE *f = fopen((pre + "_kf_frame.txt").c_str(), "w");
  for (size_t q = 0; q < m12.size(); q++) fprintf(f, "%ld\n", m12[q] ? static_cast<long>(m12[q] - store.data()) : -1L);
  fclose(f);
  f = fopen((pre + "_kps_a.txt").c_str(), "w");
  for (size_t i = 0; i < ka.size(); i++) fprintf(f, "%.1f %.1f %.9g %.9g %.9g\n", ka[i].pt.x, ka[i].pt.y, ka[i].response, cov[i].x(), cov[i].y());
  fclose(f);
  {  // throughput mode of the shim: same key points, fp16-rounded descriptors, heat maps fetched on demand and bit-identical
    setenv("SPFE_SHIM_LAZY_HEAT", "1", 1);
    setenv("SPFE_SHIM_DESC_F16", "1", 1);
    SPExtractor lazy(800);
    unsetenv("SPFE_SHIM_LAZY_HEAT");
    unsetenv("SPFE_SHIM_DESC_F16");
    std::vector<cv::KeyPoint> kl;
    cv::Mat dl;
    lazy(b, cv::Mat(), kl, dl);
    bool same = kl.size() == kb.size() && lazy.heat_.empty();
    double worst = 0;
    for (size_t i = 0; same && i < kl.size(); i++) {
      same = kl[i].pt.x == kb[i].pt.x && kl[i].pt.y == kb[i].pt.y && kl[i].response == kb[i].response;
      for (int k = 0; k < 256; k++) worst = std::max(worst, (double)std::fabs(dl.at<float>((int)i, k) - db.at<float>((int)i, k)));
    }
    cv::Mat h = lazy.getHeatMap(), hi = lazy.getHeatInv();
    same = same && !h.empty() && memcmp(h.data, sp->heat_.data, (size_t)camera::height * camera::width * 4) == 0 &&
           memcmp(hi.data, sp->heat_inv_.data, (size_t)camera::height * camera::width * 4) == 0;
    printf("throughput mode (lazy heat, fp16 descriptors): %s, max |d desc| %.2e\n", same ? "identical" : "DIFFERENT", worst);
    if (!same || worst > 5e-4) return 3;
  }
  bool threw = false;
  try { std::vector<cv::KeyPoint> k; cv::Mat d; (*ex)(cv::Mat(), cv::Mat(), k, d); } catch (const std::runtime_error &e) { threw = std::string(e.what()) == "input image is empty"; }
  printf("empty image throws runtime_error(\"input image is empty\"): %s\n", threw ? "yes" : "NO");
  delete ex;
  return threw ? 0 : 1;
}
