// Native latency of spfe_search_guided_sets (descriptors resident on the device) next to the reference-shaped CPU loop
// (oracle/sp_post.c::orc_search_guided, pinned to the reference's own SearchByProjection) on the same inputs -- without
// the Python wrapper in the timed region.  Measurement tool, not product code.
//   g++ -O2 -std=c++17 -I include tools/guided_native.cc -L sp_orb_slam_b200/lib -lspfe -L oracle/_build -lsporacle \
//       -Wl,-rpath,\$ORIGIN/../../sp_orb_slam_b200/lib -Wl,-rpath,\$ORIGIN/../../oracle/_build -o tools/_build/guided_native
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "spfe.h"

extern "C" void orc_search_guided(int m, const float *qdesc, const uint8_t *qvalid, const uint8_t *qblocks, const float *qxy,
                                  const float *qr, int mode, const int16_t *occ, int grid_rows, int grid_cols, const float *kp_un,
                                  const float *kdesc, int n, uint8_t *kp_taken, float min_x, float min_y, float best_init,
                                  float th_le, float th_lt, float c2, int32_t *q2kp, float *qdist);

int main(int argc, char **argv) {
  const char *weights = argc > 1 ? argv[1] : "tests/golden/superpoint_v1.spw";
  const int H = 480, W = 752, hc = H / 8, wc = W / 8;
  spfe_config cfg;
  spfe_default_config(&cfg, H, W, 800);
  cfg.weights_path = weights;
  cfg.flags = 0;
  spfe_ctx *ctx = nullptr;
  if (spfe_create(&cfg, &ctx) != SPFE_OK) { fprintf(stderr, "%s\n", spfe_last_error(nullptr)); return 1; }
  std::mt19937 rng(7);
  std::normal_distribution<float> gauss(0.f, 1.f);
  std::uniform_real_distribution<float> uni(0.f, 1.f);
  // a frame: one key point in ~13 % of the cells (710 of 5640), unit descriptors, raster order
  std::vector<int16_t> occ(hc * wc, -1);
  std::vector<float> kp, kdesc;
  int n = 0;
  for (int cy = 1; cy < hc - 1; cy++)
    for (int cx = 1; cx < wc - 1 && n < 710; cx++)
      if (uni(rng) < 0.1345f) {
        occ[cy * wc + cx] = (int16_t)n++;
        kp.push_back(cx * 8 + (int)(uni(rng) * 8));
        kp.push_back(cy * 8 + (int)(uni(rng) * 8));
        float ss = 0;
        size_t o = kdesc.size();
        for (int k = 0; k < 256; k++) { kdesc.push_back(gauss(rng)); ss += kdesc.back() * kdesc.back(); }
        for (int k = 0; k < 256; k++) kdesc[o + k] /= std::sqrt(ss);
      }
  for (int m : {300, 1000}) {
    const float r = m == 300 ? 4.0f : 7.0f;
    std::vector<float> qdesc((size_t)m * 256), qxy(m * 2), qr(m, r);
    for (int i = 0; i < m; i++) {
      const int src = rng() % n;
      float ss = 0;
      for (int k = 0; k < 256; k++) { qdesc[(size_t)i * 256 + k] = kdesc[(size_t)src * 256 + k] + 0.02f * gauss(rng); ss += qdesc[(size_t)i * 256 + k] * qdesc[(size_t)i * 256 + k]; }
      for (int k = 0; k < 256; k++) qdesc[(size_t)i * 256 + k] /= std::sqrt(ss);
      qxy[2 * i] = kp[2 * src] + (uni(rng) * 6 - 3);
      qxy[2 * i + 1] = kp[2 * src + 1] + (uni(rng) * 6 - 3);
    }
    spfe_desc_set *qs = nullptr, *ks = nullptr;
    spfe_desc_set_create(ctx, 4096, &qs);
    spfe_desc_set_create(ctx, 1024, &ks);
    spfe_desc_set_upload(ctx, qs, qdesc.data(), m);
    spfe_desc_set_upload(ctx, ks, kdesc.data(), n);
    spfe_guided_search g;
    memset(&g, 0, sizeof g);
    g.struct_size = sizeof g; g.mode = SPFE_GUIDED_AREA; g.m = m; g.n = n;
    g.qxy = qxy.data(); g.qradius = qr.data(); g.kp_un = kp.data(); g.occ_grid = occ.data(); g.grid_rows = hc; g.grid_cols = wc;
    g.best_init = 256.f; g.th_le = 0.7f; g.th_lt = 0.7f;
    std::vector<int32_t> q2kp(m), ref(m);
    std::vector<float> qd(m), rd(m);
    for (int i = 0; i < 20; i++)
      if (spfe_search_guided_sets(ctx, &g, qs, ks, q2kp.data(), qd.data(), nullptr) != SPFE_OK) { fprintf(stderr, "%s\n", spfe_last_error(ctx)); return 1; }
    const int reps = 500;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) spfe_search_guided_sets(ctx, &g, qs, ks, q2kp.data(), qd.data(), nullptr);
    const double t_set = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
    g.qdesc = qdesc.data(); g.kdesc = kdesc.data();
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 100; i++) spfe_search_guided(ctx, &g, q2kp.data(), qd.data(), nullptr);
    const double t_host = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 100;
    std::vector<uint8_t> taken(n);
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 100; i++) {
      memset(taken.data(), 0, n);
      orc_search_guided(m, qdesc.data(), nullptr, nullptr, qxy.data(), qr.data(), 0, occ.data(), hc, wc, kp.data(), kdesc.data(), n, taken.data(),
                        0.f, 0.f, 256.f, 0.7f, 0.7f, 0.f, ref.data(), rd.data());
    }
    const double t_cpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 100;
    int same = 1, matched = 0;
    for (int i = 0; i < m; i++) { same &= q2kp[i] == ref[i]; matched += ref[i] >= 0; }
    printf("native m=%d n=%d r=%.0f: spfe_search_guided_sets %.1f us, spfe_search_guided (host descriptors) %.1f us, CPU loop %.1f us "
           "(%.2fx), identical %d, matches %d\n", m, n, r, t_set * 1e6, t_host * 1e6, t_cpu * 1e6, t_cpu / t_set, same, matched);
    spfe_desc_set_destroy(ctx, qs);
    spfe_desc_set_destroy(ctx, ks);
  }
  spfe_destroy(ctx);
  return 0;
}
