/* TEST-ONLY stand-in for the one libspfe entry that cpp/optimizer_dust.h calls, so that the shim's host-side logic
 * (gathering map points, Converter::toSE3Quat / toCvMat, is_visible / is_mp_visible_ / dust_proj write-back) can run
 * in the CPU test suite: spfe_dust_pose_optimize is answered by the oracle (oracle/dust_pose.c).  Never linked into
 * the product; the real entry is covered by the -m gpu tests. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "spfe.h"

typedef struct { const float *dust; int rows, cols; double fx, fy, cx, cy, huber; } orc_dust_cam;
int orc_dust_optimize(const orc_dust_cam *c, double *pose7, const double *Xw, int n, int iterations, double chi2_inlier,
                      uint8_t *level, double *err, float *uv, uint8_t *visible, int *n_inlier, double *stats);

int spfe_dust_pose_optimize(spfe_ctx *ctx, const spfe_dust_pose *p, double *pose7, uint8_t *visible, float *proj_uv,
                            int32_t *n_inlier, int32_t *n_iter, double *stats) {
  (void)ctx;
  if (!p || p->struct_size != (int32_t)sizeof(spfe_dust_pose) || !p->dust) return SPFE_ERR_INVALID;
  const int n = p->n, nn = n > 0 ? n : 1;
  orc_dust_cam cam = {p->dust, p->rows, p->cols, p->fx, p->fy, p->cx, p->cy, p->huber_delta};
  uint8_t *level = (uint8_t *)calloc(nn, 1), *vis = (uint8_t *)calloc(nn, 1);
  double *err = (double *)calloc(nn, sizeof(double));
  float *uv = (float *)calloc(2 * nn, sizeof(float));
  int inl = 0;
  const int it = orc_dust_optimize(&cam, pose7, p->Xw, n, p->iterations, p->chi2_inlier, level, err, uv, vis, &inl, stats);
  if (visible) memcpy(visible, vis, n);
  if (proj_uv) memcpy(proj_uv, uv, 2 * (size_t)n * sizeof(float));
  if (n_inlier) *n_inlier = inl;
  if (n_iter) *n_iter = it;
  free(level); free(vis); free(err); free(uv);
  return it < 0 ? SPFE_ERR_STATE : SPFE_OK;
}

const char *spfe_last_error(const spfe_ctx *ctx) { (void)ctx; return "fake backend"; }
