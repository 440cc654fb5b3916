// Host check of the per-edge / per-step device functions of csrc/dustpose.cuh against oracle/dust_pose.c -- a
// development tool (no GPU in the build container): the header's __device__ functions are compiled for the host with
// the round-to-nearest intrinsics mapped to plain IEEE operations (-ffp-contract=off), driven by a sequential
// re-enactment of the kernel's control flow, and compared with the oracle.  Not part of the product or the tests.
//   g++ -O2 -ffp-contract=off -I sp_orb_slam_b200/csrc tools/dustpose_hostcheck.cc oracle/dust_pose.c -o /tmp/dp_check && /tmp/dp_check
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#define DP_HOST_CHECK 1
#define __device__
#define __forceinline__ inline
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::max;
using std::min;
#include "dustpose.cuh"

extern "C" {
typedef struct { const float *dust; int rows, cols; double fx, fy, cx, cy, huber; } orc_dust_cam;
int orc_dust_linearize(const orc_dust_cam *c, const double *pose7, const double *Xw, int n, uint8_t *level, double *err, float *uv, double *J, double *Hb);
int orc_dust_optimize(const orc_dust_cam *c, double *pose7, const double *Xw, int n, int iterations, double chi2_inlier, uint8_t *level,
                      double *err, float *uv, uint8_t *visible, int *n_inlier, double *stats);
}
using namespace spfe;

int main() {
  std::mt19937 rng(7);
  std::uniform_real_distribution<double> U(0, 1);
  int bad = 0;
  for (int trial = 0; trial < 40; trial++) {
    const int rows = 60, cols = 94, n = 50 + trial * 7;
    std::vector<float> dust(rows * cols);
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols; x++) dust[y * cols + x] = (float)(0.55 + 0.4 * std::sin(0.31 * x + 0.1 * trial) * std::cos(0.23 * y) + 0.05 * U(rng));
    double pose[7] = {0.01 * (U(rng) - .5), 0.01 * (U(rng) - .5), 0.01 * (U(rng) - .5), 1, 0.05 * (U(rng) - .5), 0.05 * (U(rng) - .5), 0.05 * (U(rng) - .5)};
    double nr = std::sqrt(pose[0] * pose[0] + pose[1] * pose[1] + pose[2] * pose[2] + 1);
    for (int k = 0; k < 4; k++) pose[k] /= nr;
    DustPoseArgs a{};
    a.dust = dust.data(); a.rows = rows; a.cols = cols; a.fx = 458.0 / 8; a.fy = 457.0 / 8; a.cx = (367.0 - 3.5) / 8; a.cy = (248.0 - 3.5) / 8;
    a.huber = 0.9; a.chi2_inlier = 0.9; a.n = n;
    std::vector<double> Xw(3 * n);
    for (int i = 0; i < n; i++) {
      const double z = 1.0 + 6 * U(rng), u = -8 + (cols + 16) * U(rng), v = -6 + (rows + 12) * U(rng);
      Xw[3 * i] = (u - a.cx) / a.fx * z; Xw[3 * i + 1] = (v - a.cy) / a.fy * z; Xw[3 * i + 2] = (trial % 5 == 0 && i % 17 == 0) ? -z : z;
    }
    orc_dust_cam cam{dust.data(), rows, cols, a.fx, a.fy, a.cx, a.cy, a.huber};
    // linearize: bit-exact per edge
    std::vector<uint8_t> lv0(n, 0), lv1(n, 0);
    std::vector<double> e0(n), e1(n), J0(6 * n), J1(6 * n), Hb(43);
    std::vector<float> uv0(2 * n, 0.f), uv1(2 * n, 0.f);
    orc_dust_linearize(&cam, pose, Xw.data(), n, lv0.data(), e0.data(), uv0.data(), J0.data(), Hb.data());
    DpPose p;
    memcpy(p.q, pose, 32); memcpy(p.t, pose + 4, 24);
    for (int i = 0; i < n; i++) {
      e1[i] = dp_error(a, dust.data(), p, &Xw[3 * i], lv1[i], &uv1[2 * i]);
      dp_jacobian(a, dust.data(), p, &Xw[3 * i], lv1[i], &J1[6 * i]);
    }
    if (memcmp(e0.data(), e1.data(), 8 * n) || memcmp(J0.data(), J1.data(), 48 * n) || lv0 != lv1 || memcmp(uv0.data(), uv1.data(), 8 * n)) { printf("trial %d: per-edge mismatch\n", trial); bad++; }
    // optimise: sequential re-enactment of the kernel
    double po[7]; memcpy(po, pose, sizeof po);
    std::vector<uint8_t> vis0(n), lv(n); int inl0 = 0; double stats[3];
    const int it0 = orc_dust_optimize(&cam, po, Xw.data(), n, 40, 0.9, lv.data(), e0.data(), uv0.data(), vis0.data(), &inl0, stats);
    DpPose pose_s = p, backup; double H[36], b[6], x[6] = {0}, lambda = 0, ni = 2, cur = 0; int ok = 1, it = 0, trials = 0;
    std::fill(lv1.begin(), lv1.end(), 0); std::fill(uv1.begin(), uv1.end(), 0.f);
    auto errors = [&]() { double chi = 0; for (int i = 0; i < n; i++) { e1[i] = dp_error(a, dust.data(), pose_s, &Xw[3 * i], lv1[i], &uv1[2 * i]); double r0, r1; dp_huber(a.huber, e1[i] * e1[i], r0, r1); chi += r0; } return chi; };
    for (; it < 40; it++) {
      if (!ok) break;
      cur = errors();
      memset(H, 0, sizeof H); memset(b, 0, sizeof b);
      for (int i = 0; i < n; i++) {
        double J[6], r0, r1; dp_jacobian(a, dust.data(), pose_s, &Xw[3 * i], lv1[i], J); dp_huber(a.huber, e1[i] * e1[i], r0, r1);
        for (int r = 0; r < 6; r++) { b[r] += J[r] * (-e1[i] * r1); for (int s = 0; s < 6; s++) H[6 * r + s] += J[r] * r1 * J[s]; }
      }
      if (it == 0) { double md = 0; for (int j = 0; j < 6; j++) md = fmax(fabs(H[7 * j]), md); lambda = 1e-5 * md; ni = 2; }
      double rho = 0; int qmax = 0; bool again;
      do {
        backup = pose_s;
        bool ok2 = dp_chol_solve6(H, lambda, b, x);
        DpPose e, np; dp_se3_exp(x, e); dp_se3_mul(e, pose_s, np); pose_s = np;
        double tmp = errors(); if (!ok2) tmp = DBL_MAX;
        rho = cur - tmp; double scale = 0; for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]); scale += 1e-3; rho /= scale;
        bool brk = false;
        if (rho > 0 && std::isfinite(tmp)) { double alpha = 1. - pow(2 * rho - 1, 3.0); alpha = fmin(alpha, 2. / 3.); lambda *= fmax(1. / 3., alpha); ni = 2; cur = tmp; }
        else { lambda *= ni; ni *= 2; pose_s = backup; if (!std::isfinite(lambda)) brk = true; }
        if (!brk) { qmax++; trials++; }
        again = !brk && rho < 0 && qmax < 10;
        if (!again && (qmax == 10 || rho == 0 || !std::isfinite(lambda))) ok = 0;
      } while (again);
    }
    double dmax = 0; for (int k = 0; k < 4; k++) dmax = fmax(dmax, fabs(pose_s.q[k] - po[k])); for (int k = 0; k < 3; k++) dmax = fmax(dmax, fabs(pose_s.t[k] - po[4 + k]));
    int inl1 = 0, visdiff = 0; for (int i = 0; i < n; i++) { bool badp = lv1[i] == 1 || e1[i] * e1[i] > 0.9; inl1 += !badp; visdiff += (vis0[i] != (uint8_t)!badp); }
    printf("trial %2d n %3d: it %d/%d trials %d/%d inliers %d/%d visdiff %d pose diff %.2e lambda %.3e/%.3e chi %.6f\n", trial, n, it0, it, (int)stats[2], trials, inl0, inl1, visdiff, dmax, stats[0], lambda, stats[1]);
    if (it0 != it || inl0 != inl1 || visdiff || dmax > 1e-11) bad++;   // the device functions use reciprocals where the oracle divides
  }
  printf(bad ? "FAILED %d\n" : "all ok\n", bad);
  return bad != 0;
}
