/*
 * CPU oracle, dust-map pose optimisation -- TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py).  Plain-C, double-precision restatement of
 *
 *   orc_dust_error      <- EdgeSE3ProjectDustOnlyPose::computeError     orb_slam2/src/optimization/types_dust_tracking.cpp:62-94
 *                          (isInImage :36-41, getPixelValue :43-56)
 *   orc_dust_jacobian   <- EdgeSE3ProjectDustOnlyPose::linearizeOplus   :96-141
 *   orc_dust_linearize  <- one computeActiveErrors + buildSystem pass of the graph that
 *                          Optimizer::PoseOptimizationDust(Frame*, mps, is_visible) builds
 *                          orb_slam2/src/mapping/optimizer_dust.cpp:170-293 (one SE3 vertex, one unary edge per map point,
 *                          identity information, Huber delta 0.9)
 *   orc_dust_optimize   <- optimizer.optimize(40) + the inlier read-out of the same function (:246-265)
 *
 * The first two follow the reference's own source expression by expression (evaluation order kept, float where the
 * reference is float, no FMA contraction: build with -ffp-contract=off; the reference is built -O3 -march=native,
 * CMakeLists.txt:8, so its own last bits depend on the build host).
 *
 * The rest lives in un-vendored third-party code: g2o (deps/g2o_catkin, .spslam_https.install:13-15, tracking upstream
 * master, no pinned revision) and Eigen 3.3.  Restated here from their published algorithms:
 *   Eigen  Quaternion * Vector3  (uv = 2 q.vec x v; v + w uv + q.vec x uv), Quaternion(Matrix3) (Shepperd),
 *   g2o    SE3Quat::map / operator* / exp / normalizeRotation, VertexSE3Expmap::oplusImpl (left-multiplied exp),
 *          RobustKernelHuber::robustify, BaseUnaryEdge::constructQuadraticForm (weightedOmega = rho[1] Omega,
 *          b += J^T (-rho[1] Omega e)), OptimizationAlgorithmLevenberg::solve (tau 1e-5, good-step scale in
 *          [1/3, 2/3], ni doubling, 10 trials, rho==0 terminates), SparseOptimizer::optimize, push / pop of the vertex
 *          estimate, LinearSolverDense (here: Cholesky of H + lambda I; g2o uses Eigen::LDLT -- same solution to
 *          rounding).
 * Parity status: the reference ships no test or golden vector for this path.  The per-edge half (orc_dust_error,
 * orc_dust_jacobian) IS pinned: oracle/ref_build.sh compiles the reference's own EdgeSE3ProjectDustOnlyPose verbatim
 * (types_dust_tracking.h:22-65, types_dust_tracking.cpp:36-141) against a g2o / Eigen stand-in (oracle/ref_g2o_stub.h)
 * into oracle/_ref/libspdust_ref.so, and tests/test_pose_dust.py::test_reference_edge_pins_oracle finds error, level,
 * (u_, v_) and Jacobian bit-identical.  The Levenberg loop (orc_dust_optimize) is PARITY UNPINNED: g2o cannot be built
 * here.
 *
 * Behaviour kept on purpose: setLevel(1) is sticky (an edge that left the image once keeps a zero Jacobian but its
 * error still enters chi2 when it projects inside again); u_ / v_ and _error are those of the LAST computeError call,
 * including rejected LM trials (pop() restores only the vertex); linearizeOplus throws when its own projection falls
 * outside the image (reported as return code -1).  Where the reference would read outside the dust map (float
 * rounding of u + 1 at the right / bottom edge, undefined behaviour) the index is clamped.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  double q[4]; /* x y z w */
  double t[3];
} orc_pose;

typedef struct {
  const float *dust;
  int rows, cols;
  double fx, fy, cx, cy;
  double huber; /* <= 0: no robust kernel */
} orc_dust_cam;

/* Eigen: Quaternion::_transformVector, then g2o SE3Quat::map = _r * xyz + _t */
static void pose_map(const orc_pose *p, const double *v, double *o) {
  const double qx = p->q[0], qy = p->q[1], qz = p->q[2], qw = p->q[3];
  double uv0 = qy * v[2] - qz * v[1], uv1 = qz * v[0] - qx * v[2], uv2 = qx * v[1] - qy * v[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  const double c0 = qy * uv2 - qz * uv1, c1 = qz * uv0 - qx * uv2, c2 = qx * uv1 - qy * uv0;
  o[0] = ((v[0] + qw * uv0) + c0) + p->t[0];
  o[1] = ((v[1] + qw * uv1) + c1) + p->t[1];
  o[2] = ((v[2] + qw * uv2) + c2) + p->t[2];
}

static int in_image(const orc_dust_cam *c, double u, double v) { /* border = 1.0; w_, h_ are floats */
  const double border = 1.0, w = (double)(float)c->cols, h = (double)(float)c->rows;
  return u >= border && u + border + 1 < w && v >= border && v + border + 1 < h;
}

static float dust_at(const orc_dust_cam *c, int y, int x) {
  if (x < 0) x = 0; if (x > c->cols - 1) x = c->cols - 1;
  if (y < 0) y = 0; if (y > c->rows - 1) y = c->rows - 1;
  return c->dust[(size_t)y * c->cols + x];
}

static float pixel_value(const orc_dust_cam *c, float x, float y) {
  const int x_f = (int)floorf(x), y_f = (int)floorf(y);
  const float xx = x - (float)x_f, yy = y - (float)y_f;
  return (1 - xx) * (1 - yy) * dust_at(c, y_f, x_f) + xx * (1 - yy) * dust_at(c, y_f, x_f + 1) +
         (1 - xx) * yy * dust_at(c, y_f + 1, x_f) + xx * yy * dust_at(c, y_f + 1, x_f + 1);
}

/* computeError: returns _error(0,0); *level |= 1 when behind the camera / outside; uv updated only when inside */
double orc_dust_error(const orc_dust_cam *c, const orc_pose *p, const double *Xw, uint8_t *level, float *uv) {
  double xl[3];
  pose_map(p, Xw, xl);
  if (xl[2] < 0.0) { *level = 1; return 0.0; }
  const double x = xl[0] * c->fx / xl[2] + c->cx, y = xl[1] * c->fy / xl[2] + c->cy;
  if (!in_image(c, x, y)) { *level = 1; return 0.0; }
  uv[0] = (float)x; uv[1] = (float)y;
  return (double)pixel_value(c, (float)x, (float)y);
}

/* linearizeOplus: J[6]; returns 0, or -1 where the reference throws std::runtime_error(" should be omitted") */
int orc_dust_jacobian(const orc_dust_cam *c, const orc_pose *p, const double *Xw, uint8_t level, double *J) {
  if (level == 1) { memset(J, 0, 6 * sizeof(double)); return 0; }
  double xl[3];
  pose_map(p, Xw, xl);
  const double x = xl[0], y = xl[1], invz = 1.0 / xl[2], invz_2 = invz * invz;
  const double u = x * c->fx * invz + c->cx, v = y * c->fy * invz + c->cy;
  if (!in_image(c, u, v)) { memset(J, 0, 6 * sizeof(double)); return -1; }
  double a0[6], a1[6];
  a0[0] = -x * y * invz_2 * c->fx;
  a0[1] = (1 + (x * x * invz_2)) * c->fx;
  a0[2] = -y * invz * c->fx;
  a0[3] = invz * c->fx;
  a0[4] = 0;
  a0[5] = -x * invz_2 * c->fx;
  a1[0] = -(1 + y * y * invz_2) * c->fy;
  a1[1] = x * y * invz_2 * c->fy;
  a1[2] = x * invz * c->fy;
  a1[3] = 0;
  a1[4] = invz * c->fy;
  a1[5] = -y * invz_2 * c->fy;
  const double g0 = (double)((pixel_value(c, (float)(u + 1), (float)v) - pixel_value(c, (float)(u - 1), (float)v)) / 2.0f);
  const double g1 = (double)((pixel_value(c, (float)u, (float)(v + 1)) - pixel_value(c, (float)u, (float)(v - 1))) / 2.0f);
  for (int k = 0; k < 6; k++) J[k] = g0 * a0[k] + g1 * a1[k];
  return 0;
}

/* g2o RobustKernelHuber::robustify */
static void huber(double delta, double e2, double *rho) {
  if (delta <= 0.0) { rho[0] = e2; rho[1] = 1.0; rho[2] = 0.0; return; }
  const double dsqr = delta * delta;
  if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.0; rho[2] = 0.0; }
  else {
    const double sqrte = sqrt(e2);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e2;
  }
}

static double compute_errors(const orc_dust_cam *c, const orc_pose *p, const double *Xw, int n, uint8_t *level, double *err,
                             float *uv) { /* computeActiveErrors + activeRobustChi2 */
  double chi = 0.0;
  for (int i = 0; i < n; i++) {
    err[i] = orc_dust_error(c, p, Xw + 3 * i, level + i, uv + 2 * i);
    double rho[3];
    huber(c->huber, err[i] * err[i], rho);
    chi += rho[0];
  }
  return chi;
}

static int build_system(const orc_dust_cam *c, const orc_pose *p, const double *Xw, int n, const uint8_t *level,
                        const double *err, double *Jout, double *H, double *b) {
  int thrown = 0;
  memset(H, 0, 36 * sizeof(double));
  memset(b, 0, 6 * sizeof(double));
  for (int i = 0; i < n; i++) {
    double Jl[6], *J = Jout ? Jout + 6 * i : Jl, rho[3];
    if (orc_dust_jacobian(c, p, Xw + 3 * i, level[i], J)) thrown = -1;
    huber(c->huber, err[i] * err[i], rho);
    const double omega_r = -err[i] * rho[1];
    for (int r = 0; r < 6; r++) {
      b[r] += J[r] * omega_r;
      for (int s = 0; s < 6; s++) H[6 * r + s] += J[r] * rho[1] * J[s];
    }
  }
  return thrown;
}

/* one computeActiveErrors + buildSystem at a fixed pose.  Hb = H[36] row-major, b[6], robust chi2. */
int orc_dust_linearize(const orc_dust_cam *c, const double *pose7, const double *Xw, int n, uint8_t *level, double *err,
                       float *uv, double *J, double *Hb) {
  orc_pose p;
  memcpy(p.q, pose7, 4 * sizeof(double));
  memcpy(p.t, pose7 + 4, 3 * sizeof(double));
  Hb[42] = compute_errors(c, &p, Xw, n, level, err, uv);
  return build_system(c, &p, Xw, n, level, err, J, Hb, Hb + 36);
}

/* ---- g2o SE3Quat pieces ---- */
static void quat_normalize_rotation(double *q) { /* SE3Quat::normalizeRotation */
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= nrm; q[1] /= nrm; q[2] /= nrm; q[3] /= nrm;
}

static void quat_from_matrix(const double R[3][3], double *q) { /* Eigen quaternionbase_assign_impl<Matrix3> */
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

static void se3_exp(const double *upd, orc_pose *out) { /* SE3Quat::exp: upd = (omega, upsilon) */
  const double *w = upd, *ups = upd + 3;
  const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double O[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
  double O2[3][3], R[3][3], V[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
  double ra, rb, va, vb;
  if (theta < 0.00001) { ra = 1.0; rb = 0.5; va = 0.5; vb = 1.0 / 6.0; }
  else {
    ra = sin(theta) / theta;
    rb = (1 - cos(theta)) / (theta * theta);
    va = rb;
    vb = (theta - sin(theta)) / (theta * theta * theta);
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const double I = i == j ? 1.0 : 0.0;
      R[i][j] = I + ra * O[i][j] + rb * O2[i][j];
      V[i][j] = I + va * O[i][j] + vb * O2[i][j];
    }
  quat_from_matrix(R, out->q);
  quat_normalize_rotation(out->q);
  for (int i = 0; i < 3; i++) out->t[i] = V[i][0] * ups[0] + V[i][1] * ups[1] + V[i][2] * ups[2];
}

static void se3_mul(const orc_pose *a, const orc_pose *b, orc_pose *o) { /* SE3Quat::operator* */
  const double ax = a->q[0], ay = a->q[1], az = a->q[2], aw = a->q[3], bx = b->q[0], by = b->q[1], bz = b->q[2], bw = b->q[3];
  orc_pose rot = *a, r;
  rot.t[0] = rot.t[1] = rot.t[2] = 0.0;
  pose_map(&rot, b->t, r.t); /* _r * tr2._t */
  r.t[0] += a->t[0]; r.t[1] += a->t[1]; r.t[2] += a->t[2];
  r.q[3] = aw * bw - ax * bx - ay * by - az * bz; /* Eigen quat product */
  r.q[0] = aw * bx + ax * bw + ay * bz - az * by;
  r.q[1] = aw * by + ay * bw + az * bx - ax * bz;
  r.q[2] = aw * bz + az * bw + ax * by - ay * bx;
  quat_normalize_rotation(r.q);
  *o = r;
}

static int chol_solve6(const double *H, double lambda, const double *b, double *x) { /* (H + lambda I) x = b */
  double L[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j <= i; j++) {
      double s = H[6 * i + j] + (i == j ? lambda : 0.0);
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return 0;
        L[i][i] = sqrt(s);
      } else
        L[i][j] = s / L[j][j];
    }
  double y[6];
  for (int i = 0; i < 6; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
  return 1;
}

/* optimizer.optimize(iterations) on the pose-only dust graph, then the inlier read-out (optimizer_dust.cpp:246-265).
 * pose7 in / out.  visible[i] = !(level == 1 || chi2 > chi2_inlier).  Returns the iterations run (g2o's return
 * value), or -1 where linearizeOplus throws.  stats (may be NULL): [0] final lambda, [1] final robust chi2 of the
 * accepted state, [2] total LM trials. */
int orc_dust_optimize(const orc_dust_cam *c, double *pose7, const double *Xw, int n, int iterations, double chi2_inlier,
                      uint8_t *level, double *err, float *uv, uint8_t *visible, int *n_inlier, double *stats) {
  orc_pose p;
  memcpy(p.q, pose7, 4 * sizeof(double));
  memcpy(p.t, pose7 + 4, 3 * sizeof(double));
  double H[36], b[6], x[6] = {0, 0, 0, 0, 0, 0}, lambda = 0.0, ni = 2.0, cur = 0.0;
  int it = 0, ok = 1, trials = 0;
  memset(level, 0, n);
  for (int i = 0; i < n; i++) { err[i] = 0.0; uv[2 * i] = uv[2 * i + 1] = 0.0f; }
  for (; it < iterations && ok; it++) {
    cur = compute_errors(c, &p, Xw, n, level, err, uv);
    if (build_system(c, &p, Xw, n, level, err, 0, H, b)) return -1;
    if (it == 0) {
      double md = 0.0;
      for (int j = 0; j < 6; j++) md = fmax(fabs(H[7 * j]), md);
      lambda = 1e-5 * md;
      ni = 2.0;
    }
    double rho = 0.0;
    int qmax = 0;
    do {
      const orc_pose backup = p; /* push */
      const int ok2 = chol_solve6(H, lambda, b, x);
      orc_pose e, np;
      se3_exp(x, &e);
      se3_mul(&e, &p, &np);
      p = np;
      double tmp = compute_errors(c, &p, Xw, n, level, err, uv);
      if (!ok2) tmp = DBL_MAX;
      rho = cur - tmp;
      double scale = 0.0;
      for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tmp)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = fmin(alpha, 2. / 3.);
        const double sf = fmax(1. / 3., alpha);
        lambda *= sf;
        ni = 2;
        cur = tmp;
      } else {
        lambda *= ni;
        ni *= 2;
        p = backup; /* pop */
        if (!isfinite(lambda)) break;
      }
      qmax++;
      trials++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0 || !isfinite(lambda)) ok = 0; /* Terminate: this iteration still counts */
  }
  int inl = n;
  for (int i = 0; i < n; i++) {
    const int bad = level[i] == 1 || err[i] * err[i] > chi2_inlier;
    visible[i] = !bad;
    if (bad) inl--;
  }
  *n_inlier = inl;
  memcpy(pose7, p.q, 4 * sizeof(double));
  memcpy(pose7 + 4, p.t, 3 * sizeof(double));
  if (stats) { stats[0] = lambda; stats[1] = cur; stats[2] = (double)trials; }
  return it;
}
