// CPU run of the guided-search templates of sp_orb_slam_b200/cpp/sp_matcher.h (SearchByProjection(Frame&, MapPoints),
// DustAssociate) against the fake backend (fake_spfe_guided.c = the oracle behind spfe_search_guided).  The pytest
// compares the resulting Frame::mvpMapPoints with the REFERENCE's own functions (oracle/_ref/libspguided_ref.so).
// usage: matcher_shim_cpu <scene.bin> <out.txt>
// scene.bin: int32 m, n, grid_rows, grid_cols; float th, th_dist, c2; float qdesc[m*256], qxy[m*2], quv[m*2], view_cos[m];
//            uint8 in_view[m], bad[m]; int32 nobs[m]; float kdesc[n*256], kp_un[n*2]; int16 occ[rows*cols]; uint8 kp_taken[n]
#include <cstdio>
#include <fstream>
#include <iostream>

#include "sp_matcher.h"

using namespace orbslam;

struct MapPoint {
  bool bad = false, mbTrackInView = true, in_view = true, dust_match = false;
  float mTrackProjX = 0, mTrackProjY = 0, mTrackViewCos = 1.f, dust_proj_u = 0, dust_proj_v = 0;
  int mnTrackScaleLevel = 0, nobs = 1;
  cv::Mat desc;
  bool isBad() const { return bad; }
  int Observations() const { return nobs; }
  cv::Mat getDescTrack() const { return desc; }
};
struct KeyFrame {
  cv::Mat mDescriptors;
  std::vector<MapPoint *> mps;
  std::vector<MapPoint *> GetMapPointMatches() { return mps; }
};
struct Frame {
  cv::Mat mDescriptors, occ_grid;
  int N = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<MapPoint *> mvpMapPoints;
  std::vector<float> mvScaleFactors{1.0f};
};

template <class T> static std::vector<T> rd(std::ifstream &f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char *>(v.data()), static_cast<std::streamsize>(n * sizeof(T)));
  return v;
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  auto hdr = rd<int32_t>(f, 4);
  const int m = hdr[0], n = hdr[1], rows = hdr[2], cols = hdr[3];
  auto par = rd<float>(f, 3);
  auto qdesc = rd<float>(f, 256 * (size_t)m), qxy = rd<float>(f, 2 * (size_t)m), quv = rd<float>(f, 2 * (size_t)m), cosv = rd<float>(f, m);
  auto in_view = rd<uint8_t>(f, m), bad = rd<uint8_t>(f, m);
  auto nobs = rd<int32_t>(f, m);
  auto kdesc = rd<float>(f, 256 * (size_t)n), kp_un = rd<float>(f, 2 * (size_t)n);
  auto occ = rd<int16_t>(f, (size_t)rows * cols);
  auto taken = rd<uint8_t>(f, n);
  if (!f) { std::cerr << "short scene file\n"; return 2; }
  SPMatcher::SetBackend(reinterpret_cast<spfe_ctx *>(0x1));
  std::vector<MapPoint> mps(m);
  std::vector<MapPoint *> vp;
  for (int i = 0; i < m; i++) {
    mps[i].desc.create(1, 256, CV_32FC1); memcpy(mps[i].desc.data, &qdesc[256 * (size_t)i], 1024);
    mps[i].mTrackProjX = qxy[2 * i]; mps[i].mTrackProjY = qxy[2 * i + 1]; mps[i].mTrackViewCos = cosv[i];
    mps[i].dust_proj_u = quv[2 * i]; mps[i].dust_proj_v = quv[2 * i + 1];
    mps[i].mbTrackInView = mps[i].in_view = in_view[i] != 0; mps[i].bad = bad[i] != 0; mps[i].nobs = nobs[i];
    vp.push_back(&mps[i]);
  }
  MapPoint holder; holder.nobs = 1;
  auto make_frame = [&](bool with_taken) {
    Frame F; F.N = n;
    F.mDescriptors.create(n, 256, CV_32FC1); memcpy(F.mDescriptors.data, kdesc.data(), kdesc.size() * 4);
    F.occ_grid.create(rows, cols, CV_16SC1); memcpy(F.occ_grid.data, occ.data(), occ.size() * 2);
    for (int k = 0; k < n; k++) F.mvKeysUn.push_back(cv::KeyPoint(kp_un[2 * k], kp_un[2 * k + 1], 1.0f));
    F.mvpMapPoints.assign(n, nullptr);
    if (with_taken) for (int k = 0; k < n; k++) if (taken[k]) F.mvpMapPoints[k] = &holder;
    return F;
  };
  SPMatcher matcher(0.7f);
  FILE *o = fopen(argv[2], "w");
  {
    Frame F = make_frame(true);
    const int nm = matcher.SearchByProjection(F, vp, par[0], par[1], par[2]);
    fprintf(o, "%d\n", nm);
    for (int k = 0; k < n; k++) fprintf(o, "%ld ", (F.mvpMapPoints[k] && F.mvpMapPoints[k] != &holder) ? (long)(F.mvpMapPoints[k] - mps.data()) : -1L);
    fprintf(o, "\n");
  }
  {
    Frame F = make_frame(false);
    const int nm = matcher.DustAssociate(F, vp);
    fprintf(o, "%d\n", nm);
    for (int k = 0; k < n; k++) fprintf(o, "%ld ", F.mvpMapPoints[k] ? (long)(F.mvpMapPoints[k] - mps.data()) : -1L);
    fprintf(o, "\n");
    for (int i = 0; i < m; i++) fprintf(o, "%d", mps[i].dust_match ? 1 : 0);
    fprintf(o, "\n");
  }
  {  // SearchByBruteForce: key frame = the map points' descriptors (rows without / with bad map points), frame = the keypoints
    KeyFrame kf, kf2;
    kf.mDescriptors.create(m, 256, CV_32FC1); memcpy(kf.mDescriptors.data, qdesc.data(), qdesc.size() * 4);
    for (int i = 0; i < m; i++) kf.mps.push_back(in_view[i] ? &mps[i] : nullptr);   // in_view reused as "has a map point"
    Frame F = make_frame(false);
    std::vector<MapPoint *> m12;
    matcher.SearchByBruteForce(&kf, F, m12);
    for (int q = 0; q < n; q++) fprintf(o, "%ld ", m12[q] ? (long)(m12[q] - mps.data()) : -1L);
    fprintf(o, "\n");
    std::vector<MapPoint> store2(n);
    kf2.mDescriptors = F.mDescriptors;
    for (int k = 0; k < n; k++) kf2.mps.push_back(taken[k] ? nullptr : &store2[k]);         // taken reused as "no map point"
    std::vector<MapPoint *> mkk;
    const int cnt = matcher.SearchByBruteForce(&kf, &kf2, mkk);
    fprintf(o, "%d\n", cnt);
    for (int i = 0; i < m; i++) fprintf(o, "%ld ", mkk[i] ? (long)(mkk[i] - store2.data()) : -1L);
    fprintf(o, "\n");
  }
  fclose(o);
  return 0;
}
