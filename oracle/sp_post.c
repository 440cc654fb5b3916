/*
 * CPU oracle, post-processing half -- TEST INFRASTRUCTURE ONLY (see
 * oracle/__init__.py).  Plain-C restatement of the reference's host-side
 * algorithms with cv::Mat / Eigen replaced by flat arrays.  Parity status:
 * unpinned by the reference (it ships no tests); pinned here against the
 * reference's OWN nms() / computeCovariance() compiled verbatim into
 * oracle/_ref/libsppost_ref.so (oracle/ref_build.sh, oracle/ref_cv_stub.h;
 * bit-identical in tests/test_oracle.py::test_reference_*_pins_oracle) and
 * against OpenCV (cv2.BFMatcher / sortIdx) for the third-party calls.
 *
 *   orc_to_heat       <- orb_slam2/src/cv/sp_extractor.cpp:461-474 (to_heat lambda)
 *   orc_sort_desc     <- :489-498 (cv::sortIdx, SORT_DESCENDING; ties: lower index first)
 *   orc_nms           <- :161-250 (nms)
 *   orc_covariance    <- :252-340 (computeCovariance)
 *   orc_l2            <- orb_slam2/src/cv/sp_matcher.cpp:1636-1640 (DescriptorDistance)
 *   orc_match_mutual  <- :1666-1669 and sp_matcher_loop.cpp:365-368
 *                        (cv::BFMatcher(NORM_L2, crossCheck=true)::match)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* heat = (-x - min)/(max - min), heat_inv = (max - (-x))/(max - min), evaluated the
 * way OpenCV's MatExpr folds them: one convertTo(alpha, beta) per output with
 * alpha/beta computed in double, narrowed to float, applied as x*alpha + beta. */
void orc_to_heat(const float *heat_log, int n, float *heat, float *heat_inv,
                 double *min_out, double *max_out) {
  double mn = DBL_MAX, mx = -DBL_MAX;
  for (int i = 0; i < n; i++) {
    double v = (double)(heat_log[i] * -1.0f);
    if (v < mn) mn = v;
    if (v > mx) mx = v;
  }
  const double inv = 1.0 / (mx - mn);
  const float a0 = (float)(-1.0 * inv), b0 = (float)((-mn) * inv);
  const float a1 = (float)(1.0 * inv), b1 = (float)(mx * inv);
  for (int i = 0; i < n; i++) {
    volatile float p0 = heat_log[i] * a0; /* volatile: forbid fma contraction */
    volatile float p1 = heat_log[i] * a1;
    heat[i] = p0 + b0;
    heat_inv[i] = p1 + b1;
  }
  if (min_out) *min_out = mn;
  if (max_out) *max_out = mx;
}

typedef struct { float s; int i; } si_t;
static int cmp_desc(const void *a, const void *b) {
  const si_t *x = (const si_t *)a, *y = (const si_t *)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}
void orc_sort_desc(const float *score, int n, int32_t *order) {
  si_t *v = (si_t *)malloc(sizeof(si_t) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) { v[i].s = score[i]; v[i].i = i; }
  qsort(v, n, sizeof(si_t), cmp_desc);
  for (int i = 0; i < n; i++) order[i] = v[i].i;
  free(v);
}

/* Greedy radius-`r` suppression over candidates given in descending score
 * order (pts = n x 2 floats, x then y).  Returns the number of survivors N;
 * sel[N] = indices into the sorted candidate list in raster (v outer, u inner)
 * order; occ[hc*wc] = survivor index per 8x8 cell or -1. */
int orc_nms(const float *pts, int n, int num_features, int border, int r,
            int W, int H, int32_t *sel, int16_t *occ) {
  const int PW = W + 2 * r, PH = H + 2 * r;
  uint8_t *grid = (uint8_t *)calloc((size_t)PW * PH, 1);
  uint16_t *inds = (uint16_t *)calloc((size_t)W * H, sizeof(uint16_t));
  for (int i = 0; i < (W / 8) * (H / 8); i++) occ[i] = -1;
  for (int i = 0; i < n; i++) {
    int u = (int)pts[2 * i], v = (int)pts[2 * i + 1];
    grid[(size_t)(v + r) * PW + (u + r)] = 1;
    inds[(size_t)v * W + u] = (uint16_t)i;
  }
  int kept = 0;
  for (int i = 0; i < n; i++) {
    int u = (int)pts[2 * i] + r, v = (int)pts[2 * i + 1] + r;
    if (grid[(size_t)v * PW + u] != 1) continue;
    for (int k = -r; k <= r; k++)
      for (int j = -r; j <= r; j++)
        if (j || k) grid[(size_t)(v + k) * PW + (u + j)] = 0;
    grid[(size_t)v * PW + u] = 2;
    if (++kept > num_features) break;
  }
  int N = 0;
  for (int v = border; v < H - border; v++)
    for (int u = border; u < W - border; u++)
      if (grid[(size_t)(v + r) * PW + (u + r)] == 2) {
        occ[(v / 8) * (W / 8) + (u / 8)] = (int16_t)N;
        sel[N++] = inds[(size_t)v * W + u];
      }
  free(grid);
  free(inds);
  return N;
}

/* Flood-fill second moments around each keypoint on heat_inv (h x w).
 * kps = N x 2 floats (x, y) in output order; the visited map is shared by all
 * keypoints and a pixel is marked when popped, exactly as in the reference. */
void orc_covariance(const float *heat, int h, int w, const float *kps, int N,
                    float *response, float *cov2, float *cov2_inv) {
  uint8_t *fresh = (uint8_t *)malloc((size_t)h * w);
  memset(fresh, 1, (size_t)h * w);
  size_t cap = 1 << 16, head, tail;
  int32_t *q = (int32_t *)malloc(cap * sizeof(int32_t));
  size_t pcap = 1 << 16, np_;
  float *du = (float *)malloc(pcap * sizeof(float));
  float *dv = (float *)malloc(pcap * sizeof(float));
  float *sc = (float *)malloc(pcap * sizeof(float));
  for (int k = 0; k < N; k++) {
    const int uu = (int)kps[2 * k], vv = (int)kps[2 * k + 1];
    response[k] = heat[(size_t)vv * w + uu];
    head = tail = 0; np_ = 0;
    q[tail++] = vv * w + uu;
    while (head < tail) {
      const int p = q[head++];
      const int u = p % w, v = p / w;
      fresh[p] = 0;
      if (np_ == pcap) {
        pcap *= 2;
        du = (float *)realloc(du, pcap * sizeof(float));
        dv = (float *)realloc(dv, pcap * sizeof(float));
        sc = (float *)realloc(sc, pcap * sizeof(float));
      }
      const float fu = (float)u - (float)uu, fv = (float)v - (float)vv;
      du[np_] = fu * fu; dv[np_] = fv * fv; sc[np_] = heat[p]; np_++;
      const float centroid = heat[p];
      int nb[4], m = 0;
      if (u - 1 > 0) nb[m++] = p - 1;      /* left  */
      if (v - 1 > 0) nb[m++] = p - w;      /* up    */
      if (u + 1 < w) nb[m++] = p + 1;      /* right */
      if (v + 1 < h) nb[m++] = p + w;      /* down  */
      for (int t = 0; t < m; t++) {
        const float hv = heat[nb[t]];
        if (fresh[nb[t]] && hv > 0.0f && hv < centroid) {
          if (tail == cap) { cap *= 2; q = (int32_t *)realloc(q, cap * sizeof(int32_t)); }
          q[tail++] = nb[t];
        }
      }
    }
    float sum = 0.0f;
    for (size_t i = 0; i < np_; i++) sum += sc[i];
    float cx = 0.0f, cy = 0.0f;
    for (size_t i = 0; i < np_; i++) {
      const float wgt = sc[i] / sum;
      cx += wgt * du[i];
      cy += wgt * dv[i];
    }
    if (cx < 1.0f) cx = 1.0f;
    if (cy < 1.0f) cy = 1.0f;
    cov2[2 * k] = cx; cov2[2 * k + 1] = cy;
    cov2_inv[2 * k] = 1.0f / cx; cov2_inv[2 * k + 1] = 1.0f / cy;
  }
  free(fresh); free(q); free(du); free(dv); free(sc);
}

float orc_l2(const float *a, const float *b, int d) {
  float s = 0.0f;
  for (int i = 0; i < d; i++) { const float t = a[i] - b[i]; s += t * t; }
  return sqrtf(s);
}

/* BFMatcher(NORM_L2, crossCheck=true).match(query) against one train set:
 * q2t[i] = first-index arg-min train row of query i if that train row's
 * first-index arg-min query row is i, else -1.  dist[i] = that L2 distance.
 * second[i] (optional) = 2nd-smallest distance of query i (parity margins). */
void orc_match_mutual(const float *q, int nq, const float *t, int nt, int d,
                      int32_t *q2t, float *dist, float *second) {
  int32_t *t2q = (int32_t *)malloc(sizeof(int32_t) * (nt > 0 ? nt : 1));
  float *tbest = (float *)malloc(sizeof(float) * (nt > 0 ? nt : 1));
  for (int j = 0; j < nt; j++) { t2q[j] = -1; tbest[j] = FLT_MAX; }
  for (int i = 0; i < nq; i++) {
    float best = FLT_MAX, sec = FLT_MAX; int bj = -1;
    for (int j = 0; j < nt; j++) {
      const float dd = orc_l2(q + (size_t)i * d, t + (size_t)j * d, d);
      if (dd < best) { sec = best; best = dd; bj = j; }
      else if (dd < sec) sec = dd;
      if (dd < tbest[j]) { tbest[j] = dd; t2q[j] = i; }
    }
    q2t[i] = bj; dist[i] = best;
    if (second) second[i] = sec;
  }
  for (int i = 0; i < nq; i++)
    if (q2t[i] < 0 || t2q[q2t[i]] != i) q2t[i] = -1;
  free(t2q); free(tbest);
}

/* ---------------------------------------------------------------------------
 * Guided (cell-grid) searches -- SURVEY.md section 8(f) rank 2.
 *
 *   orc_features_in_area <- Frame::GetFeaturesInArea, orb_slam2/src/type/frame.cpp:382-420: the live code walks the
 *                           extractor's occ_grid (8-px cells, ix outer / iy inner) and keeps the keypoints with
 *                           |dx| < r && |dy| < r; minLevel / maxLevel are ignored by it.
 *   orc_search_guided    <- the greedy loops of
 *        SPMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, th_dist)  sp_matcher.cpp:344-432   (mode 0)
 *        SPMatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono)       sp_matcher.cpp:1439-1543 (mode 0)
 *        the dust-track patch association of Tracking                                  tracker_dust.cpp:112-172 (mode 1)
 *   with the per-object tests hoisted into flat arrays: qvalid[i] = the loop's `continue` tests passed
 *   (mbTrackInView && !isBad(), pMP && !mvbOutlier && projected inside the image, in_view && !isBad()), qblocks[i] =
 *   the matched keypoint becomes unavailable to later queries (pMP->Observations() > 0; always for mode 1, which
 *   clears the occ_grid cell), kp_taken[idx] = the keypoint already carries an observed map point on entry.
 *   Acceptance: best <= th_le  ||  best < (c2 > 0 ? 1.2f * c2 / (c2 + duv) : th_lt).
 *   Pinned against the reference's own GetFeaturesInArea / SearchByProjection(Frame&, MapPoints) / association block,
 *   compiled verbatim into oracle/_ref/libspguided_ref.so (tests/test_guided.py::test_reference_*_pins_oracle).
 *   The reference dereferences mvKeysUn[-1] when every candidate of a map point is skipped (sp_matcher.cpp:416-417,
 *   undefined behaviour); here that is "no match".  Mode 1 reads occ_grid without a bounds check upstream; here
 *   out-of-range cells are skipped.
 * ------------------------------------------------------------------------- */
int orc_features_in_area(const int16_t *occ, int grid_rows, int grid_cols, const float *kp_un, float x, float y, float r,
                         float min_x, float min_y, int32_t *out) {
  int n = 0;
  int c0 = (int)floorf((x - min_x - r) / 8.0f); if (c0 < 0) c0 = 0;
  if (c0 >= grid_cols) return 0;
  int c1 = (int)ceilf((x - min_x + r) / 8.0f); if (c1 > grid_cols - 1) c1 = grid_cols - 1;
  if (c1 < 0) return 0;
  int r0 = (int)floorf((y - min_y - r) / 8.0f); if (r0 < 0) r0 = 0;
  if (r0 >= grid_rows) return 0;
  int r1 = (int)ceilf((y - min_y + r) / 8.0f); if (r1 > grid_rows - 1) r1 = grid_rows - 1;
  if (r1 < 0) return 0;
  for (int ix = c0; ix <= c1; ix++)
    for (int iy = r0; iy <= r1; iy++) {
      const int16_t idx = occ[iy * grid_cols + ix];
      if (idx == -1) continue;
      const float dx = kp_un[2 * idx] - x, dy = kp_un[2 * idx + 1] - y;
      if (fabsf(dx) < r && fabsf(dy) < r) out[n++] = idx;
    }
  return n;
}

void orc_search_guided(int m, const float *qdesc, const uint8_t *qvalid, const uint8_t *qblocks, const float *qxy,
                       const float *qr, int mode, const int16_t *occ, int grid_rows, int grid_cols, const float *kp_un,
                       const float *kdesc, int n, uint8_t *kp_taken, float min_x, float min_y, float best_init,
                       float th_le, float th_lt, float c2, int32_t *q2kp, float *qdist) {
  int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(grid_rows * grid_cols + 4));
  (void)n;
  for (int i = 0; i < m; i++) {
    q2kp[i] = -1;
    qdist[i] = 0.0f;
    if (qvalid && !qvalid[i]) continue;
    const float x = qxy[2 * i], y = qxy[2 * i + 1];
    int nc = 0;
    if (mode == 0) {
      nc = orc_features_in_area(occ, grid_rows, grid_cols, kp_un, x, y, qr[i], min_x, min_y, cand);
    } else {
      const int u = (int)floorf(x), v = (int)floorf(y);
      for (int du = 0; du < 2; du++)
        for (int dv = 0; dv < 2; dv++) {
          const int uu = u + du, vv = v + dv;
          if (uu < 0 || uu >= grid_cols || vv < 0 || vv >= grid_rows) continue;
          const int16_t idx = occ[vv * grid_cols + uu];
          if (idx != -1) cand[nc++] = idx;
        }
    }
    float best = best_init;
    int best_idx = -1;
    for (int c = 0; c < nc; c++) {
      const int idx = cand[c];
      if (kp_taken[idx]) continue;
      const float d = orc_l2(qdesc + (size_t)i * 256, kdesc + (size_t)idx * 256, 256);
      if (d < best) { best = d; best_idx = idx; }
    }
    if (best_idx < 0) continue;
    int accept = best <= th_le;
    if (!accept) {
      float thr = th_lt;
      if (c2 > 0.0f) {
        const float du = kp_un[2 * best_idx] - x, dv = kp_un[2 * best_idx + 1] - y;
        const float duv = du * du + dv * dv;
        thr = 1.2f * c2 / (c2 + duv);
      }
      accept = best < thr;
    }
    if (!accept) continue;
    q2kp[i] = best_idx;
    qdist[i] = best;
    if (!qblocks || qblocks[i]) kp_taken[best_idx] = 1;
  }
  free(cand);
}

/* Exact 2 nearest neighbours (first index on ties): the result FLANN's KD-tree search approximates in
 * SPMatcher::SearchForTriByFlann / SearchByFlann (sp_matcher.cpp:197-206, :266-270); == cv::BFMatcher(NORM_L2).knnMatch(q, t, 2). */
void orc_knn2(const float *q, int nq, const float *t, int nt, int d, int32_t *idx, float *dist) {
  for (int i = 0; i < nq; i++) {
    float b0 = FLT_MAX, b1 = FLT_MAX;
    int j0 = -1, j1 = -1;
    for (int j = 0; j < nt; j++) {
      const float dd = orc_l2(q + (size_t)i * d, t + (size_t)j * d, d);
      if (dd < b0) { b1 = b0; j1 = j0; b0 = dd; j0 = j; }
      else if (dd < b1) { b1 = dd; j1 = j; }
    }
    idx[2 * i] = j0; idx[2 * i + 1] = j1;
    dist[2 * i] = j0 >= 0 ? b0 : 0.0f; dist[2 * i + 1] = j1 >= 0 ? b1 : 0.0f;
  }
}
