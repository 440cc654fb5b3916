// Dust-map pose optimisation on the device -- SURVEY.md section 8(f) rank 4.
//
// Replaces the inner loop of Optimizer::PoseOptimizationDust(Frame*, mps, is_visible)
// (orb_slam2/src/mapping/optimizer_dust.cpp:170-293): a g2o graph with one SE3 vertex and one unary edge per map point,
// EdgeSE3ProjectDustOnlyPose (orb_slam2/src/optimization/types_dust_tracking.cpp:36-141), whose error is the bilinear
// sample of the dustbin probability map (hc x wc, Frame::dust_ = SPExtractor::dense_dust_) at the projection of the
// point, run for 40 Levenberg-Marquardt iterations with a Huber kernel (delta 0.9).
//
//   dust_pose_kernel   ONE CTA does the whole solve in one launch: the dust map is staged in shared memory once
//                      (60 x 94 x 4 B = 22 KB; 1080p: 127 KB), every thread owns the edges i = tid, tid + 256, ...
//                      (error, sticky level, Jacobian in registers), the 6 x 6 normal equations are a fixed-order
//                      block reduction (warp shuffles, then warps in order: deterministic), and thread 0 runs the
//                      g2o Levenberg step (Cholesky of H + lambda I, SE3 exp, accept / reject, lambda schedule)
//                      between two barriers.  No host round trip per iteration: a 40-iteration solve is one launch
//                      and one 7-double copy back, and the dust map never has to leave the device
//                      (spfe_dust_pose.dust == NULL reads the slot's dense_dust in place).
//                      mode 0 = one computeActiveErrors + buildSystem pass (spfe_dust_linearize).
//
// Per-edge arithmetic follows the reference expression by expression: double where it is double, float inside
// getPixelValue, the reference's evaluation order, and round-to-nearest intrinsics so that nvcc cannot contract a
// multiply-add the host build does not (bit-exact against oracle/dust_pose.c).  The Levenberg step is g2o's published
// algorithm (see the oracle's header for the list); only the reduction order differs from g2o's edge-by-edge sums.
#pragma once
#ifndef DP_HOST_CHECK  // tools/dustpose_hostcheck.cc compiles the per-edge / per-step functions below for the host
#include <cuda_runtime.h>
#endif
#include <float.h>
#include <stdint.h>

namespace spfe {

constexpr int DP_THREADS = 256;
constexpr int DP_NRED = 28;  // 21 (upper triangle of H) + 6 (b) + 1 (edges whose linearizeOplus would throw)

struct DustPoseArgs {
  const float *dust;   // [rows][cols] (global)
  int rows, cols, dust_in_smem;
  const double *Xw;    // [n][3]
  int n, mode, iterations;
  double fx, fy, cx, cy, huber, chi2_inlier;
  const double *pose_in;   // [7] qx qy qz qw tx ty tz
  const uint8_t *level_in; // [n] or null (= all 0): levels on entry (mode 0)
  double *pose;        // [7] out (mode 1)
  uint8_t *level;      // [n] out
  double *err;         // [n]
  float *uv;           // [n][2]
  double *J;           // [n][6]   (mode 0 only)
  double *Hb;          // [43]     (mode 0: H, b, chi2;  mode 1: [0] lambda, [1] chi2 of the accepted state, [2] trials)
  uint8_t *visible;    // [n]      (mode 1)
  int *result;         // [0] iterations run, or -1 if linearizeOplus would have thrown; [1] inliers
};

struct DpPose { double q[4], t[3]; };

#define DP_MUL(a, b) __dmul_rn((a), (b))
#define DP_ADD(a, b) __dadd_rn((a), (b))
#define DP_SUB(a, b) __dsub_rn((a), (b))
#define DP_DIV(a, b) __ddiv_rn((a), (b))
#define DP_FMUL(a, b) __fmul_rn((a), (b))
#define DP_FADD(a, b) __fadd_rn((a), (b))
#define DP_FSUB(a, b) __fsub_rn((a), (b))

// Eigen Quaternion * Vector3 (uv = 2 q.vec x v; v + w uv + q.vec x uv), then + t (g2o SE3Quat::map)
__device__ __forceinline__ void dp_map(const DpPose &p, const double *v, double *o) {
  const double qx = p.q[0], qy = p.q[1], qz = p.q[2], qw = p.q[3];
  double uv0 = DP_SUB(DP_MUL(qy, v[2]), DP_MUL(qz, v[1]));
  double uv1 = DP_SUB(DP_MUL(qz, v[0]), DP_MUL(qx, v[2]));
  double uv2 = DP_SUB(DP_MUL(qx, v[1]), DP_MUL(qy, v[0]));
  uv0 = DP_ADD(uv0, uv0); uv1 = DP_ADD(uv1, uv1); uv2 = DP_ADD(uv2, uv2);
  const double c0 = DP_SUB(DP_MUL(qy, uv2), DP_MUL(qz, uv1));
  const double c1 = DP_SUB(DP_MUL(qz, uv0), DP_MUL(qx, uv2));
  const double c2 = DP_SUB(DP_MUL(qx, uv1), DP_MUL(qy, uv0));
  o[0] = DP_ADD(DP_ADD(DP_ADD(v[0], DP_MUL(qw, uv0)), c0), p.t[0]);
  o[1] = DP_ADD(DP_ADD(DP_ADD(v[1], DP_MUL(qw, uv1)), c1), p.t[1]);
  o[2] = DP_ADD(DP_ADD(DP_ADD(v[2], DP_MUL(qw, uv2)), c2), p.t[2]);
}

// isInImage, border = 1.0 (types_dust_tracking.cpp:36-41); w_, h_ are floats there
__device__ __forceinline__ bool dp_in_image(const DustPoseArgs &a, double u, double v) {
  const double w = (double)(float)a.cols, h = (double)(float)a.rows;
  return u >= 1.0 && DP_ADD(DP_ADD(u, 1.0), 1.0) < w && v >= 1.0 && DP_ADD(DP_ADD(v, 1.0), 1.0) < h;
}

__device__ __forceinline__ float dp_at(const DustPoseArgs &a, const float *dust, int y, int x) {
  x = min(max(x, 0), a.cols - 1);  // the reference reads out of bounds here (float rounding at the far edge); clamp
  y = min(max(y, 0), a.rows - 1);
  return dust[y * a.cols + x];
}

// getPixelValue (types_dust_tracking.cpp:43-56), all float, left to right
__device__ __forceinline__ float dp_pixel(const DustPoseArgs &a, const float *dust, float x, float y) {
  const int x_f = (int)floorf(x), y_f = (int)floorf(y);
  const float xx = DP_FSUB(x, (float)x_f), yy = DP_FSUB(y, (float)y_f);
  const float ox = DP_FSUB(1.0f, xx), oy = DP_FSUB(1.0f, yy);
  const float t0 = DP_FMUL(DP_FMUL(ox, oy), dp_at(a, dust, y_f, x_f));
  const float t1 = DP_FMUL(DP_FMUL(xx, oy), dp_at(a, dust, y_f, x_f + 1));
  const float t2 = DP_FMUL(DP_FMUL(ox, yy), dp_at(a, dust, y_f + 1, x_f));
  const float t3 = DP_FMUL(DP_FMUL(xx, yy), dp_at(a, dust, y_f + 1, x_f + 1));
  return DP_FADD(DP_FADD(DP_FADD(t0, t1), t2), t3);
}

// computeError (types_dust_tracking.cpp:62-94)
__device__ __forceinline__ double dp_error(const DustPoseArgs &a, const float *dust, const DpPose &p, const double *Xw,
                                           uint8_t &level, float *uv) {
  double xl[3];
  dp_map(p, Xw, xl);
  if (xl[2] < 0.0) { level = 1; return 0.0; }
  const double x = DP_ADD(DP_DIV(DP_MUL(xl[0], a.fx), xl[2]), a.cx);
  const double y = DP_ADD(DP_DIV(DP_MUL(xl[1], a.fy), xl[2]), a.cy);
  if (!dp_in_image(a, x, y)) { level = 1; return 0.0; }
  uv[0] = (float)x; uv[1] = (float)y;
  return (double)dp_pixel(a, dust, (float)x, (float)y);
}

// linearizeOplus (types_dust_tracking.cpp:96-141); false where the reference throws " should be omitted"
__device__ __forceinline__ bool dp_jacobian(const DustPoseArgs &a, const float *dust, const DpPose &p, const double *Xw,
                                            uint8_t level, double *J) {
#pragma unroll
  for (int k = 0; k < 6; k++) J[k] = 0.0;
  if (level == 1) return true;
  double xl[3];
  dp_map(p, Xw, xl);
  const double x = xl[0], y = xl[1], invz = DP_DIV(1.0, xl[2]), invz_2 = DP_MUL(invz, invz);
  const double u = DP_ADD(DP_MUL(DP_MUL(x, a.fx), invz), a.cx), v = DP_ADD(DP_MUL(DP_MUL(y, a.fy), invz), a.cy);
  if (!dp_in_image(a, u, v)) return false;
  double a0[6], a1[6];
  a0[0] = DP_MUL(DP_MUL(DP_MUL(-x, y), invz_2), a.fx);
  a0[1] = DP_MUL(DP_ADD(1.0, DP_MUL(DP_MUL(x, x), invz_2)), a.fx);
  a0[2] = DP_MUL(DP_MUL(-y, invz), a.fx);
  a0[3] = DP_MUL(invz, a.fx);
  a0[4] = 0.0;
  a0[5] = DP_MUL(DP_MUL(-x, invz_2), a.fx);
  a1[0] = DP_MUL(-DP_ADD(1.0, DP_MUL(DP_MUL(y, y), invz_2)), a.fy);
  a1[1] = DP_MUL(DP_MUL(DP_MUL(x, y), invz_2), a.fy);
  a1[2] = DP_MUL(DP_MUL(x, invz), a.fy);
  a1[3] = 0.0;
  a1[4] = DP_MUL(invz, a.fy);
  a1[5] = DP_MUL(DP_MUL(-y, invz_2), a.fy);
  const float fu = (float)u, fv = (float)v;
  const double g0 = (double)__fdiv_rn(DP_FSUB(dp_pixel(a, dust, (float)DP_ADD(u, 1.0), fv), dp_pixel(a, dust, (float)DP_SUB(u, 1.0), fv)), 2.0f);
  const double g1 = (double)__fdiv_rn(DP_FSUB(dp_pixel(a, dust, fu, (float)DP_ADD(v, 1.0)), dp_pixel(a, dust, fu, (float)DP_SUB(v, 1.0))), 2.0f);
#pragma unroll
  for (int k = 0; k < 6; k++) J[k] = DP_ADD(DP_MUL(g0, a0[k]), DP_MUL(g1, a1[k]));
  return true;
}

// g2o RobustKernelHuber::robustify: rho0 (robust chi2) and rho1 (weight)
__device__ __forceinline__ void dp_huber(double delta, double e2, double &rho0, double &rho1) {
  rho0 = e2;
  rho1 = 1.0;
  if (delta > 0.0 && e2 > DP_MUL(delta, delta)) {
    const double sqrte = sqrt(e2);
    rho0 = DP_SUB(DP_MUL(DP_MUL(2.0, sqrte), delta), DP_MUL(delta, delta));
    rho1 = DP_DIV(delta, sqrte);
  }
}

// ---- thread-0 pieces: g2o SE3Quat (exp, product, normalizeRotation), dense 6 x 6 solve ----
// (the serial fp64 section paces the kernel -- every other warp waits at the barrier behind it -- so divisions are
// replaced by one reciprocal where g2o / Eigen divide element by element: same value to rounding)
__device__ inline void dp_normalize_rotation(double *q) {
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double inv = rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}

__device__ inline void dp_quat_from_matrix(const double R[3][3], double *q) {  // Eigen, Shepperd
  double t = R[0][0] + R[1][1] + R[2][2];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[2][1] - R[1][2]) * t;
    q[1] = (R[0][2] - R[2][0]) * t;
    q[2] = (R[1][0] - R[0][1]) * t;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k][j] - R[j][k]) * t;
    q[j] = (R[j][i] + R[i][j]) * t;
    q[k] = (R[k][i] + R[i][k]) * t;
  }
}

__device__ inline void dp_se3_exp(const double *upd, DpPose &out) {  // upd = (omega, upsilon)
  const double *w = upd, *ups = upd + 3;
  const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double O[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
  double O2[3][3], R[3][3], V[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
  double ra, rb, va, vb;
  if (theta < 0.00001) { ra = 1.0; rb = 0.5; va = 0.5; vb = 1.0 / 6.0; }
  else {
    double sn, cs;
    sincos(theta, &sn, &cs);
    const double it = 1.0 / theta, it2 = it * it;
    ra = sn * it;
    rb = (1 - cs) * it2;
    va = rb;
    vb = (theta - sn) * (it2 * it);
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double I = i == j ? 1.0 : 0.0;
      R[i][j] = I + ra * O[i][j] + rb * O2[i][j];
      V[i][j] = I + va * O[i][j] + vb * O2[i][j];
    }
  dp_quat_from_matrix(R, out.q);
  dp_normalize_rotation(out.q);
  for (int i = 0; i < 3; i++) out.t[i] = V[i][0] * ups[0] + V[i][1] * ups[1] + V[i][2] * ups[2];
}

__device__ inline void dp_rotate(const double *q, const double *v, double *o) {
  double uv0 = q[1] * v[2] - q[2] * v[1], uv1 = q[2] * v[0] - q[0] * v[2], uv2 = q[0] * v[1] - q[1] * v[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  o[0] = v[0] + q[3] * uv0 + (q[1] * uv2 - q[2] * uv1);
  o[1] = v[1] + q[3] * uv1 + (q[2] * uv0 - q[0] * uv2);
  o[2] = v[2] + q[3] * uv2 + (q[0] * uv1 - q[1] * uv0);
}

__device__ inline void dp_se3_mul(const DpPose &a, const DpPose &b, DpPose &o) {  // SE3Quat::operator*
  DpPose r;
  dp_rotate(a.q, b.t, r.t);
  r.t[0] += a.t[0]; r.t[1] += a.t[1]; r.t[2] += a.t[2];
  const double ax = a.q[0], ay = a.q[1], az = a.q[2], aw = a.q[3], bx = b.q[0], by = b.q[1], bz = b.q[2], bw = b.q[3];
  r.q[3] = aw * bw - ax * bx - ay * by - az * bz;
  r.q[0] = aw * bx + ax * bw + ay * bz - az * by;
  r.q[1] = aw * by + ay * bw + az * bx - ax * bz;
  r.q[2] = aw * bz + az * bw + ax * by - ay * bx;
  dp_normalize_rotation(r.q);
  o = r;
}

__device__ inline bool dp_chol_solve6(const double *H, double lambda, const double *b, double *x) {  // (H + lambda I) x = b
  double L[6][6], inv[6];   // inv[j] = 1 / L[j][j]
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double s = H[6 * j + j] + lambda;
#pragma unroll
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
    if (!(s > 0.0)) return false;
    inv[j] = rsqrt(s);
    L[j][j] = s * inv[j];
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double t = H[6 * i + j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t * inv[j];
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k];
    x[i] = s * inv[i];
  }
  return true;
}

#ifndef DP_HOST_CHECK
// Fixed-order block sum of NV doubles per thread: lanes by butterfly, then the warps in index order.  Result in out[].
template <int NV>
__device__ __forceinline__ void dp_block_sum(double (&v)[NV], double (*s_part)[DP_NRED], double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_part[warp][k] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < DP_THREADS / 32; w++) s += s_part[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

struct DpState {
  DpPose pose, backup;
  double H[36], b[6], x[6];
  double red[DP_NRED];
  double lambda, ni, cur, rho;
  int qmax, again, ok, thrown, trials;
};

// computeActiveErrors + activeRobustChi2 -> st.red[0]
__device__ __forceinline__ void dp_errors_pass(const DustPoseArgs &a, const float *dust, DpState &st,
                                               double (*s_part)[DP_NRED]) {
  const DpPose p = st.pose;
  double acc[1] = {0.0};
  for (int i = threadIdx.x; i < a.n; i += DP_THREADS) {
    const double X[3] = {a.Xw[3 * i], a.Xw[3 * i + 1], a.Xw[3 * i + 2]};
    uint8_t lv = a.level[i];
    float uv[2] = {a.uv[2 * i], a.uv[2 * i + 1]};
    const double e = dp_error(a, dust, p, X, lv, uv);
    a.err[i] = e;
    a.level[i] = lv;
    a.uv[2 * i] = uv[0];
    a.uv[2 * i + 1] = uv[1];
    double r0, r1;
    dp_huber(a.huber, DP_MUL(e, e), r0, r1);
    acc[0] += r0;
  }
  dp_block_sum<1>(acc, s_part, st.red);
}

// buildSystem: linearizeOplus + constructQuadraticForm of every edge -> st.H, st.b, st.thrown
__device__ __forceinline__ void dp_system_pass(const DustPoseArgs &a, const float *dust, DpState &st,
                                               double (*s_part)[DP_NRED], bool writeJ) {
  const DpPose p = st.pose;
  double acc[DP_NRED];
#pragma unroll
  for (int k = 0; k < DP_NRED; k++) acc[k] = 0.0;
  for (int i = threadIdx.x; i < a.n; i += DP_THREADS) {
    const double X[3] = {a.Xw[3 * i], a.Xw[3 * i + 1], a.Xw[3 * i + 2]};
    double J[6];
    if (!dp_jacobian(a, dust, p, X, a.level[i], J)) acc[27] += 1.0;
    if (writeJ) {
#pragma unroll
      for (int k = 0; k < 6; k++) a.J[6 * i + k] = J[k];
    }
    const double e = a.err[i];
    double r0, r1;
    dp_huber(a.huber, DP_MUL(e, e), r0, r1);
    const double omega_r = DP_MUL(-e, r1);
    int t = 0;
#pragma unroll
    for (int r = 0; r < 6; r++) {
      const double jw = DP_MUL(J[r], r1);
#pragma unroll
      for (int s = r; s < 6; s++) acc[t++] += DP_MUL(jw, J[s]);
      acc[21 + r] += DP_MUL(J[r], omega_r);
    }
  }
  dp_block_sum<DP_NRED>(acc, s_part, st.red);
  if (threadIdx.x == 0) {
    int t = 0;
    for (int r = 0; r < 6; r++)
      for (int s = r; s < 6; s++) { st.H[6 * r + s] = st.red[t]; st.H[6 * s + r] = st.red[t]; t++; }
    for (int r = 0; r < 6; r++) st.b[r] = st.red[21 + r];
    if (st.red[27] > 0.0) st.thrown = 1;
  }
  __syncthreads();
}

// One CTA per problem: batch[blockIdx.x] (spfe_dust_pose_optimize_batch solves many frames' poses in one launch).
__global__ void __launch_bounds__(DP_THREADS) dust_pose_kernel(const DustPoseArgs *__restrict__ batch) {
  const DustPoseArgs a = batch[blockIdx.x];
  extern __shared__ float s_dust[];
  __shared__ double s_part[DP_THREADS / 32][DP_NRED];
  __shared__ DpState st;
  const float *dust = a.dust;
  if (a.dust_in_smem) {
    for (int i = threadIdx.x; i < a.rows * a.cols; i += DP_THREADS) s_dust[i] = a.dust[i];
    dust = s_dust;
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; k++) st.pose.q[k] = a.pose_in[k];
    for (int k = 0; k < 3; k++) st.pose.t[k] = a.pose_in[4 + k];
    for (int k = 0; k < 6; k++) st.x[k] = 0.0;
    st.lambda = 0.0; st.ni = 2.0; st.cur = 0.0; st.rho = 0.0;
    st.qmax = 0; st.again = 0; st.ok = 1; st.thrown = 0; st.trials = 0;
  }
  for (int i = threadIdx.x; i < a.n; i += DP_THREADS) {
    a.level[i] = a.level_in ? a.level_in[i] : 0;
    a.err[i] = 0.0; a.uv[2 * i] = 0.0f; a.uv[2 * i + 1] = 0.0f;
  }
  __syncthreads();

  if (a.mode == 0) {  // spfe_dust_linearize
    dp_errors_pass(a, dust, st, s_part);
    const double chi = st.red[0];
    __syncthreads();
    dp_system_pass(a, dust, st, s_part, true);
    if (threadIdx.x == 0) {
      for (int k = 0; k < 36; k++) a.Hb[k] = st.H[k];
      for (int k = 0; k < 6; k++) a.Hb[36 + k] = st.b[k];
      a.Hb[42] = chi;
      a.result[0] = st.thrown ? -1 : 0;
      a.result[1] = 0;
    }
    return;
  }

  // SparseOptimizer::optimize(iterations) with OptimizationAlgorithmLevenberg::solve per iteration
  int it = 0;
  for (; it < a.iterations; it++) {
    if (!st.ok) break;   // uniform: st.ok only changes between barriers
    dp_errors_pass(a, dust, st, s_part);
    if (threadIdx.x == 0) st.cur = st.red[0];
    __syncthreads();
    dp_system_pass(a, dust, st, s_part, false);
    if (st.thrown) break;
    if (threadIdx.x == 0) {
      if (it == 0) {
        double md = 0.0;
        for (int j = 0; j < 6; j++) md = fmax(fabs(st.H[7 * j]), md);
        st.lambda = 1e-5 * md;   // computeLambdaInit: tau * max diagonal
        st.ni = 2.0;
      }
      st.rho = 0.0;
      st.qmax = 0;
    }
    __syncthreads();
    bool again;
    do {
      bool ok2 = false;
      if (threadIdx.x == 0) {
        st.backup = st.pose;  // push
        ok2 = dp_chol_solve6(st.H, st.lambda, st.b, st.x);
        DpPose e, np;
        dp_se3_exp(st.x, e);
        dp_se3_mul(e, st.pose, np);
        st.pose = np;
      }
      __syncthreads();
      dp_errors_pass(a, dust, st, s_part);
      if (threadIdx.x == 0) {
        double tmp = st.red[0];
        if (!ok2) tmp = DBL_MAX;
        double rho = st.cur - tmp, scale = 0.0;
        for (int j = 0; j < 6; j++) scale += st.x[j] * (st.lambda * st.x[j] + st.b[j]);
        scale += 1e-3;
        rho /= scale;
        bool brk = false;
        if (rho > 0 && isfinite(tmp)) {
          double alpha = 1. - pow((2 * rho - 1), 3.0);
          alpha = fmin(alpha, 2. / 3.);
          st.lambda *= fmax(1. / 3., alpha);
          st.ni = 2;
          st.cur = tmp;
        } else {
          st.lambda *= st.ni;
          st.ni *= 2;
          st.pose = st.backup;  // pop
          if (!isfinite(st.lambda)) brk = true;
        }
        if (!brk) { st.qmax++; st.trials++; }
        st.rho = rho;
        st.again = !brk && rho < 0 && st.qmax < 10;
        if (!st.again && (st.qmax == 10 || rho == 0 || !isfinite(st.lambda))) st.ok = 0;  // Terminate
      }
      __syncthreads();
      again = st.again != 0;
      __syncthreads();
    } while (again);
  }
  __syncthreads();
  // inlier read-out (optimizer_dust.cpp:250-265)
  double cnt[1] = {0.0};
  for (int i = threadIdx.x; i < a.n; i += DP_THREADS) {
    const double e = a.err[i];
    const bool bad = a.level[i] == 1 || DP_MUL(e, e) > a.chi2_inlier;
    a.visible[i] = bad ? 0 : 1;
    if (!bad) cnt[0] += 1.0;
  }
  dp_block_sum<1>(cnt, s_part, st.red);
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; k++) a.pose[k] = st.pose.q[k];
    for (int k = 0; k < 3; k++) a.pose[4 + k] = st.pose.t[k];
    a.Hb[0] = st.lambda; a.Hb[1] = st.cur; a.Hb[2] = (double)st.trials;
    a.result[0] = st.thrown ? -1 : it;
    a.result[1] = (int)st.red[0];
  }
}

#endif  // DP_HOST_CHECK

}  // namespace spfe
