/* TEST-ONLY stand-in for the libspfe entries that the guided-search templates of cpp/sp_matcher.h call, so that the
 * shim's host-side logic (hoisting the per-object tests of the reference's loops into flat arrays, applying the
 * assignments in map-point order; the row filtering / index mapping of SearchByBruteForce) can run in the CPU test suite:
 * spfe_search_guided and spfe_match_mutual_nn are answered by the oracle (oracle/sp_post.c).  Never linked into the product; the real entry is covered by the -m gpu tests. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "spfe.h"

void orc_search_guided(int m, const float *qdesc, const uint8_t *qvalid, const uint8_t *qblocks, const float *qxy, const float *qr, int mode,
                       const int16_t *occ, int grid_rows, int grid_cols, const float *kp_un, const float *kdesc, int n, uint8_t *kp_taken,
                       float min_x, float min_y, float best_init, float th_le, float th_lt, float c2, int32_t *q2kp, float *qdist);
float orc_l2(const float *a, const float *b, int d);

int spfe_search_guided(spfe_ctx *ctx, const spfe_guided_search *g, int32_t *q2kp, float *qdist, uint8_t *kp_taken_out) {
  (void)ctx;
  if (!g || g->struct_size != (int32_t)sizeof(spfe_guided_search)) return SPFE_ERR_INVALID;
  const int n = g->n > 0 ? g->n : 1;
  uint8_t *taken = (uint8_t *)calloc(n, 1);
  if (g->kp_taken) memcpy(taken, g->kp_taken, g->n);
  float *kp_un = (float *)calloc(2 * (size_t)n, sizeof(float));
  if (g->kp_un) memcpy(kp_un, g->kp_un, 2 * (size_t)g->n * sizeof(float));
  orc_search_guided(g->m, g->qdesc, g->qvalid, g->qblocks, g->qxy, g->qradius, g->mode, g->occ_grid, g->grid_rows, g->grid_cols, kp_un,
                    g->kdesc, g->n, taken, g->min_x, g->min_y, g->best_init, g->th_le, g->th_lt, g->c2_adaptive, q2kp, qdist);
  if (kp_taken_out) memcpy(kp_taken_out, taken, g->n);
  free(taken); free(kp_un);
  return SPFE_OK;
}

void orc_match_mutual(const float *q, int nq, const float *t, int nt, int d, int32_t *q2t, float *dist, float *second);

int spfe_match_mutual_nn(spfe_ctx *ctx, const float *q, int32_t nq, const float *t, int32_t nt, int32_t *q2t, float *dist) {
  (void)ctx;
  float *d = (float *)calloc(nq > 0 ? nq : 1, sizeof(float));
  orc_match_mutual(q, nq, t, nt, 256, q2t, d, NULL);
  if (dist) memcpy(dist, d, (size_t)nq * sizeof(float));
  free(d);
  return SPFE_OK;
}
float spfe_l2(const float *a, const float *b) { return orc_l2(a, b, 256); }
const char *spfe_last_error(const spfe_ctx *ctx) { (void)ctx; return "fake backend"; }

void orc_knn2(const float *q, int nq, const float *t, int nt, int d, int32_t *idx, float *dist);
int spfe_match_knn2(spfe_ctx *ctx, const float *q, int32_t nq, const float *t, int32_t nt, int32_t *idx, float *dist) {
  (void)ctx;
  orc_knn2(q, nq, t, nt, 256, idx, dist);
  return SPFE_OK;
}
