/*
 * spfe.h -- C ABI of the B200-native SuperPoint front-end (extract + match).
 *
 * This is the drop-in boundary for the hot path of HyHuang1995/sp_orb_slam.
 * The only intended caller is the C++ shim in sp_orb_slam_b200/cpp/ that
 * re-creates the reference classes `orbslam::SPExtractor` and
 * `orbslam::SPMatcher` with their exact signatures.
 * Plain pointers and sizes only: no libtorch, no OpenCV, no CUDA types.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   spfe_create            <- SPExtractor::SPExtractor(int)          orb_slam2/src/cv/sp_extractor.cpp:342-359
 *                             (+ SPFrontend ctor :23-76, torch::load of common::model_path :355)
 *   spfe_extract           <- SPExtractor::operator()                orb_slam2/src/cv/sp_extractor.cpp:361-514
 *                             = SPFrontend::forward :79-159, to_heat :461-474, sortIdx :489-498,
 *                               nms :161-250, computeCovariance :252-340
 *   spfe_frame_out fields  <- public members read by Frame::ExtractORB   orb_slam2/src/type/frame.cpp:296-314
 *                             (sp_extractor.h:61-77: semi_dust_, dense_dust_, heat_, heat_inv_, occ_grid_,
 *                              getCov(), getCov2Inv())
 *   spfe_match_mutual_nn   <- cv::BFMatcher(NORM_L2, crossCheck=true)::match as used by
 *                             SPMatcher::SearchByBruteForce             orb_slam2/src/cv/sp_matcher.cpp:1642-1674
 *                                                                       orb_slam2/src/cv/sp_matcher_loop.cpp:334-376
 *   spfe_l2                <- SPMatcher::DescriptorDistance             orb_slam2/src/cv/sp_matcher.cpp:1636-1640
 *   spfe_dust_pose_optimize <- Optimizer::PoseOptimizationDust          orb_slam2/src/mapping/optimizer_dust.cpp:170-293
 *                             (EdgeSE3ProjectDustOnlyPose, orb_slam2/src/optimization/types_dust_tracking.cpp:36-141)
 *
 * The reference is one-frame-blocking with batch size 1 (sp_extractor.cpp:70
 * "TODO: batch-size").  spfe_extract keeps that contract; spfe_submit /
 * spfe_wait add the batched, pipelined entry the B200 needs to be fed.
 *
 * Error convention: every function returns SPFE_OK (0) or a negative code and
 * never throws across the ABI; spfe_last_error() returns the message.  There
 * is NO CPU fallback: if no sm_100 device is present spfe_create fails.
 */
#ifndef SPFE_H_
#define SPFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPFE_VERSION 1
#define SPFE_DESC_DIM 256

enum {
  SPFE_OK = 0,
  SPFE_ERR_INVALID = -1,   /* bad argument (NULL, size mismatch, H/W not multiple of 8, batch too large) */
  SPFE_ERR_EMPTY = -2,     /* empty image: the shim turns this into std::runtime_error("input image is empty") */
  SPFE_ERR_WEIGHTS = -3,   /* weight file missing / malformed */
  SPFE_ERR_NO_DEVICE = -4, /* no CUDA device with compute capability 10.x */
  SPFE_ERR_CUDA = -5,      /* CUDA runtime / driver error (message has details) */
  SPFE_ERR_STATE = -6      /* wait without submit, slot busy, ... */
};

enum {
  SPFE_EMIT_HEAT = 1u << 0, /* copy heat_ (H x W f32; read by Frame::ExtractORB, frame.cpp:304) to the host */
  SPFE_EMIT_COV = 1u << 1,  /* run computeCovariance (on the device; the heat maps need not leave it): fills
                               kp_response (heat_inv at the keypoint), cov2, cov2_inv */
  SPFE_EMIT_HEAT_INV = 1u << 3, /* copy heat_inv_ (H x W f32 = 1 - heat_; a public member of SPExtractor that nothing outside
                               computeCovariance reads) to the host as well */
  SPFE_LAZY_HEAT = 1u << 4, /* throughput mode: the heat maps are computed and stay on the device (what computeCovariance
                               needs), spfe_fetch_heat copies heat_ / heat_inv_ of single frames on demand.  heat_ is
                               read on the host only by PoseOptimizationHeat (orb_slam2/src/tracking/tracker.cpp:206-224,
                               off the live path), so 1.44 MB per 752x480 frame need not cross PCIe */
  SPFE_DESC_F16 = 1u << 5,  /* descriptors cross PCIe as IEEE fp16 (spfe_frame_out.desc_f16, desc == NULL): half the bytes
                               of the dominant output; the C++ shim widens them to CV_32F (cosine to the fp32 rows
                               >= 1 - 1e-6, the parity bar is 1 - 1e-3) */
  SPFE_EXACT = 1u << 6,     /* "exact" mode: fp32-equivalent convolutions on the tensor core.  Activations and weights
                               are carried as hi + lo fp16 pairs (22 significant bits) and every product is three MMAs
                               (Ah*Wh + Ah*Wl + Al*Wh) into one fp32 accumulator; conv1a runs in fp32 on the CUDA cores.
                               Logits agree with the fp32 reference to ~2e-4 (default mode: ~7e-2), which is what
                               "identical keypoint sets after NMS" needs; costs ~3x the tensor work */
  SPFE_MATCH_PREV = 1u << 2 /* a slot is one camera stream: also match every frame against the previous frame of
                               that slot (mutual NN, all descriptors as train set -- the BFMatcher call of
                               Tracking::trackReferenceKeyFrameANN, tracker.cpp:372-417); frame 0 of a batch is
                               matched against the last frame of the slot's previous batch */
};

typedef struct spfe_ctx spfe_ctx;

typedef struct spfe_config {
  int32_t struct_size;   /* = sizeof(spfe_config) */
  int32_t height;        /* camera::height, multiple of 8   (sp_extractor.cpp:354) */
  int32_t width;         /* camera::width,  multiple of 8 */
  int32_t max_keypoints; /* tracking::num_features; up to max_keypoints+1 survive (sp_extractor.cpp:211) */
  float score_thresh;    /* 0.007f  (sp_extractor.cpp:122) */
  int32_t nms_radius;    /* 4       (sp_extractor.cpp:502); must be <= 8 */
  int32_t border;        /* 8       (sp_extractor.cpp:502) */
  int32_t device_id;     /* CUDA ordinal */
  int32_t max_batch;     /* frames per submit (>= 1) */
  int32_t num_slots;     /* independent in-flight batches (>= 1), each with its own stream + buffers */
  uint32_t flags;        /* SPFE_EMIT_* */
  const char *weights_path; /* legacy superpoint.pt (PyTorch-1.0 archive) or .spw */
} spfe_config;

/* Fills *cfg with the reference's hard-coded values for the given geometry. */
void spfe_default_config(spfe_config *cfg, int32_t height, int32_t width, int32_t max_keypoints);

/* Host-visible result of one frame.  All pointers are owned by the context
 * (pinned host memory), valid until the slot is submitted again / destroyed. */
typedef struct spfe_frame_out {
  int32_t n;               /* keypoints, raster order (v outer, u inner) */
  const float *kp_xy;      /* [n][2] (x, y), integer-valued          -> cv::KeyPoint(x, y, size=1) */
  const float *kp_score;   /* [n] softmax score of the keypoint's cell */
  const float *kp_response;/* [n] heat_inv(y, x) (kp.response, sp_extractor.cpp:271); NULL without EMIT_COV */
  const float *desc;       /* [n][256] L2-normalised                  -> CV_32FC1 n x 256 */
  const int16_t *occ_grid; /* [H/8][W/8] keypoint index or -1         -> occ_grid_  (CV_16SC1) */
  const float *dense_dust; /* [H/8][W/8] softmax dustbin probability  -> dense_dust_ */
  const float *semi_dust;  /* [H/8][W/8] raw dustbin logit            -> semi_dust_ */
  const float *heat;       /* [H][W] or NULL                          -> heat_ */
  const float *heat_inv;   /* [H][W] or NULL (SPFE_EMIT_HEAT_INV)     -> heat_inv_ */
  const float *cov2;       /* [n][2] or NULL                          -> getCov() */
  const float *cov2_inv;   /* [n][2] or NULL                          -> getCov2Inv() */
  int32_t n_prev;          /* SPFE_MATCH_PREV: keypoints of the previous frame of this stream (0 = none yet) */
  const int32_t *match_prev; /* [n] index into the previous frame's keypoints or -1; NULL without SPFE_MATCH_PREV */
  const float *match_dist; /* [n] L2 distance to the nearest previous descriptor */
  const uint16_t *desc_f16;/* [n][256] IEEE binary16 bits, L2-normalised (SPFE_DESC_F16; desc is NULL then) */
} spfe_frame_out;

int spfe_create(const spfe_config *cfg, spfe_ctx **out);
void spfe_destroy(spfe_ctx *ctx);
/* Message of the last failure on this context (ctx == NULL: last spfe_create failure). */
const char *spfe_last_error(const spfe_ctx *ctx);

/* Blocking single-frame extraction == SPExtractor::operator().  gray: H rows of
 * W bytes (CV_8UC1), row_stride in bytes.  Uses slot 0. */
int spfe_extract(spfe_ctx *ctx, const uint8_t *gray, size_t row_stride, spfe_frame_out *out);

/* Throughput mode: enqueue `batch` frames (host pointers) on `slot` and return
 * immediately; spfe_wait blocks until that slot's results are on the host and
 * fills outs[0..batch).  The H2D copies, the kernels and the D2H copies of all slots run on three in-order queues
 * chained by events, so with >= 2 slots in flight the copies of one batch hide behind the kernels of another while
 * the persistent kernels of consecutive batches never compete for SMs.  (SPFE_SLOT_STREAMS=1 in the environment
 * gives every slot one private stream instead.) */
int spfe_submit(spfe_ctx *ctx, int32_t slot, const uint8_t *const *grays, int32_t batch, size_t row_stride);
int spfe_wait(spfe_ctx *ctx, int32_t slot, spfe_frame_out *outs);
/* Only the n[b] valid descriptor rows of every frame cross PCIe (not max_keypoints + 1): spfe_wait first waits for the
 * counts and the small outputs, then copies exactly n[b] rows per frame.  spfe_last_d2h_bytes returns the bytes the
 * last spfe_wait of `slot` moved device -> host in total (bench.py reports it). */
int64_t spfe_last_d2h_bytes(const spfe_ctx *ctx, int32_t slot);
/* heat_ / heat_inv_ (H x W f32 each; either pointer may be NULL) of frame `frame` of the last batch waited for on `slot`,
 * copied on demand: needs SPFE_LAZY_HEAT, SPFE_EMIT_COV or SPFE_EMIT_HEAT* (the maps exist on the device) and is
 * bit-identical to what SPFE_EMIT_HEAT / SPFE_EMIT_HEAT_INV deliver eagerly.  Valid until the slot is submitted again.
 * Replaces the eager heat_ / heat_inv_ members of SPExtractor (sp_extractor.h:72-73) in throughput mode. */
int spfe_fetch_heat(spfe_ctx *ctx, int32_t slot, int32_t frame, float *heat, float *heat_inv);
/* Zero-staging variant of spfe_submit for a capture pipeline that already owns page-locked memory: `frames` is
 * [batch][H][W] u8, dense, and is DMA'd to the device straight from the caller's buffer (no host memcpy), so it must
 * stay valid and unmodified until spfe_wait returns for this slot.  Works with pageable memory too, only slower.
 * spfe_host_alloc / spfe_host_free hand out page-locked memory without the caller needing the CUDA headers. */
int spfe_submit_pinned(spfe_ctx *ctx, int32_t slot, const uint8_t *frames, int32_t batch);
void *spfe_host_alloc(size_t bytes);
void spfe_host_free(void *p);

/* Same pipeline on frames that already live in device memory ([batch][H][W] u8,
 * dense).  Results stay on the device; no host copies.  Asynchronous on the
 * slot's stream; spfe_slot_sync waits for it. */
int spfe_submit_device(spfe_ctx *ctx, int32_t slot, const void *d_gray, int32_t batch);
int spfe_slot_sync(spfe_ctx *ctx, int32_t slot);

/* Mutual nearest neighbour under L2 over 256-d float rows (host pointers):
 * q2t[i] = first-index arg-min train row of query i if that row's first-index
 * arg-min query is i, else -1; dist[i] = L2 distance to the arg-min (may be NULL).
 * Unit-norm rows (SuperPoint descriptors) take the tensor-core path: an fp16 distance GEMM (tcgen05) nominates three
 * columns per 256-column block, every nominee within the fp16 error bound of the row's best -- and the whole block
 * where three near-ties could hide a fourth -- is re-ranked with the exact fp32 distance, so the result is the exact
 * fp32 one (match.cuh has the argument).  Rows that are not unit vectors take an exact fp32 CUDA-core kernel.
 * Thread-safe (serialised on the context), as SearchByBruteForce runs on two threads. */
int spfe_match_mutual_nn(spfe_ctx *ctx, const float *q, int32_t nq, const float *t, int32_t nt,
                         int32_t *q2t, float *dist);
/* Exact 2 nearest neighbours under L2 of every query row among the train rows (host pointers): idx[i][0..1] = best /
 * second-best train row (first index on ties, -1 if there is none), dist[i][0..1] = their L2 distances.  This is
 * cv::DescriptorMatcher::knnMatch(query, matches, 2) as called on the key frame's FLANN index by
 * SPMatcher::SearchForTriByFlann / SearchByFlann (orb_slam2/src/cv/sp_matcher.cpp:183-200, :262-270) -- exact where the
 * reference's KD-tree search is approximate; the ratio test (0.7, :203-206) stays with the caller.  Thread-safe. */
int spfe_match_knn2(spfe_ctx *ctx, const float *q, int32_t nq, const float *t, int32_t nt, int32_t *idx, float *dist);

/* Device-resident descriptor sets.  The host-pointer entries above upload both descriptor matrices on every call; a
 * SLAM front-end matches the same rows many times (a key frame's map-point descriptors against every following frame,
 * the frame's own descriptors against several key frames), so a set keeps them in HBM (fp32 rows for the exact
 * re-rank and the guided searches + an fp16 copy for the tensor-core nomination):
 *   spfe_desc_set_upload      host rows -> set (e.g. KeyFrame::mDescriptors rows of the valid map points, once per key frame)
 *   spfe_desc_set_from_frame  rows[0..n) (NULL: all key points) of frame `frame` of the last batch waited for on `slot`,
 *                             gathered device -> device: the descriptors never cross PCIe for matching
 *   spfe_match_mutual_nn_sets / spfe_match_knn2_sets   the two matchers on sets (results as in the host-pointer forms)
 * Capacity <= 4096 rows (the matcher's limit).  All set entries are thread-safe like spfe_match_mutual_nn.  Sets belong
 * to their context: spfe_destroy frees the ones still alive (their handles are dead afterwards). */
typedef struct spfe_desc_set spfe_desc_set;
int spfe_desc_set_create(spfe_ctx *ctx, int32_t capacity, spfe_desc_set **out);
void spfe_desc_set_destroy(spfe_ctx *ctx, spfe_desc_set *set);
int32_t spfe_desc_set_size(const spfe_desc_set *set);
int spfe_desc_set_upload(spfe_ctx *ctx, spfe_desc_set *set, const float *rows, int32_t n);
int spfe_desc_set_from_frame(spfe_ctx *ctx, spfe_desc_set *set, int32_t slot, int32_t frame, const int32_t *rows, int32_t n);
int spfe_match_mutual_nn_sets(spfe_ctx *ctx, const spfe_desc_set *q, const spfe_desc_set *t, int32_t *q2t, float *dist);
int spfe_match_knn2_sets(spfe_ctx *ctx, const spfe_desc_set *q, const spfe_desc_set *t, int32_t *idx, float *dist);

/* Guided (cell-grid) searches -- the greedy candidate loops of
 *   SPMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, th_dist)   orb_slam2/src/cv/sp_matcher.cpp:344-432
 *   SPMatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono)        orb_slam2/src/cv/sp_matcher.cpp:1439-1543
 *   the patch-wise association of dust tracking                                    orb_slam2/src/tracking/tracker_dust.cpp:112-172
 * with the per-object tests hoisted into flat arrays by the caller.  For every query (map
 * point) i, in order: candidates = Frame::GetFeaturesInArea(qxy[i], qradius[i]) over occ_grid (frame.cpp:382-420;
 * SPFE_GUIDED_AREA) or the 2 x 2 occ_grid cells at floor(qxy[i]) in cell units (SPFE_GUIDED_DUST_CELLS); candidates
 * with kp_taken set are skipped; the nearest descriptor (first on ties, distances below best_init only) is accepted if
 *   dist <= th_le  ||  dist < (c2_adaptive > 0 ? 1.2f * c2_adaptive / (c2_adaptive + |kp_un - qxy|^2) : th_lt)
 * and, if qblocks[i], becomes unavailable to the queries that follow.  Results are those of the sequential loop.
 *   SearchByProjection(F, MPs):     best_init 256, th_le = th_dist, th_lt = 0.7 (or c2_adaptive = tracking::dust::c2_thresh)
 *   SearchByProjection(Cur, Last):  best_init FLT_MAX, th_le = TH_HIGH (0.7), th_lt = -INFINITY
 *   dust-track association:         best_init 0.75, th_le = -INFINITY, th_lt = 0.75, qblocks = NULL
 * q2kp[i] = matched keypoint or -1, qdist[i] = its distance; kp_taken_out (may be NULL) = kp_taken after the loop.
 * Thread-safe like spfe_match_mutual_nn. */
enum { SPFE_GUIDED_AREA = 0, SPFE_GUIDED_DUST_CELLS = 1 };
typedef struct spfe_guided_search {
  int32_t struct_size;      /* = sizeof(spfe_guided_search) */
  int32_t mode;             /* SPFE_GUIDED_* */
  int32_t m;                /* queries, in the reference's iteration order */
  int32_t n;                /* keypoints of the frame */
  const float *qdesc;       /* [m][256] MapPoint::getDescTrack() */
  const float *qxy;         /* [m][2] projection: pixels (AREA) or occ_grid cell units (DUST_CELLS) */
  const float *qradius;     /* [m] search radius in pixels (AREA); ignored for DUST_CELLS */
  const uint8_t *qvalid;    /* [m] or NULL: 0 = the reference's loop `continue`s before searching */
  const uint8_t *qblocks;   /* [m] or NULL (= all 1): pMP->Observations() > 0 */
  const float *kdesc;       /* [n][256] Frame::mDescriptors */
  const float *kp_un;       /* [n][2] Frame::mvKeysUn (x, y); may be NULL for DUST_CELLS */
  const int16_t *occ_grid;  /* [grid_rows][grid_cols] Frame::occ_grid (keypoint index or -1) */
  int32_t grid_rows, grid_cols;
  const uint8_t *kp_taken;  /* [n] or NULL: keypoint already carries an observed map point */
  float min_x, min_y;       /* Frame::mnMinX / mnMinY */
  float best_init, th_le, th_lt, c2_adaptive;
} spfe_guided_search;
int spfe_search_guided(spfe_ctx *ctx, const spfe_guided_search *g, int32_t *q2kp, float *qdist, uint8_t *kp_taken_out);
/* The same search with the descriptors already on the device: qset replaces g->qdesc (m rows: the local map's
 * getDescTrack() rows, uploaded when the local map changes), kset replaces g->kdesc (n rows: the current frame's
 * descriptors, spfe_desc_set_from_frame).  Either may be NULL (then the host pointer of g is used).  Only the small
 * per-call arrays (projections, radii, flags, occ_grid: ~20 KB) cross PCIe, in one block each way. */
int spfe_search_guided_sets(spfe_ctx *ctx, const spfe_guided_search *g, const spfe_desc_set *qset, const spfe_desc_set *kset,
                            int32_t *q2kp, float *qdist, uint8_t *kp_taken_out);

/* Dust-map pose optimisation -- the inner loop of
 *   Optimizer::PoseOptimizationDust(Frame*, const vector<MapPoint*>&, vector<bool>&)   orb_slam2/src/mapping/optimizer_dust.cpp:170-293
 * i.e. a g2o graph of one SE3 vertex and one EdgeSE3ProjectDustOnlyPose per map point
 *   computeError / linearizeOplus / isInImage / getPixelValue                          orb_slam2/src/optimization/types_dust_tracking.cpp:36-141
 * solved by 40 Levenberg iterations with a Huber kernel, in ONE kernel launch (dustpose.cuh).  The caller hoists the
 * map points into a flat array; the intrinsics are those the reference hands the edges: fx / 8, fy / 8, (cx - 3.5) / 8,
 * (cy - 3.5) / 8 (optimizer_dust.cpp:222-225).  pose = Frame::mTcw as (qx, qy, qz, qw, tx, ty, tz), g2o::SE3Quat order.
 * dust == NULL reads the dense_dust map of frame `frame` of `slot`'s last completed batch where it already lies in
 * device memory (the map then never crosses PCIe for tracking); otherwise dust = Frame::dust_ on the host.
 *   spfe_dust_pose_optimize: pose in / out; visible[i] = !(level == 1 || chi2 > chi2_inlier) (:250-265, is_visible /
 *     MapPoint::in_view); proj_uv[i] = the edge's (u_, v_) (-> MapPoint::dust_proj_u / v, valid where visible);
 *     *n_inlier = the function's return value; *n_iter = optimizer.optimize()'s.  stats (may be NULL) receives
 *     {final lambda, robust chi2 of the accepted state, LM trials}.
 *   spfe_dust_linearize: one computeActiveErrors + buildSystem pass at a fixed pose, for callers that keep g2o in
 *     charge of the iteration: level[n] in / out (setLevel(1) is sticky), err[n] = _error, proj_uv, J[n][6] =
 *     _jacobianOplusXi, Hb[43] = H (6 x 6, row-major), b (6), robust chi2.
 * Both return SPFE_ERR_STATE where linearizeOplus would throw std::runtime_error(" should be omitted") (:114-116; its
 * projection falls outside the image although computeError's did not); the shim rethrows.  Thread-safe like
 * spfe_match_mutual_nn. */
typedef struct spfe_dust_pose {
  int32_t struct_size;   /* = sizeof(spfe_dust_pose) */
  int32_t n;             /* map points, in the caller's order */
  const double *Xw;      /* [n][3] MapPoint::GetWorldPos() */
  const float *dust;     /* [rows][cols] or NULL (= device-resident dense_dust of slot / frame) */
  int32_t rows, cols;    /* H / 8, W / 8 */
  int32_t slot, frame;   /* used when dust == NULL */
  double fx, fy, cx, cy; /* in dust-map units */
  double huber_delta;    /* 0.9 (optimizer_dust.cpp:219); <= 0 = no robust kernel */
  double chi2_inlier;    /* 0.9 (optimizer_dust.cpp:253) */
  int32_t iterations;    /* 40  (optimizer_dust.cpp:246) */
  int32_t reserved;
} spfe_dust_pose;
int spfe_dust_pose_optimize(spfe_ctx *ctx, const spfe_dust_pose *p, double *pose7, uint8_t *visible, float *proj_uv,
                            int32_t *n_inlier, int32_t *n_iter, double *stats);
/* `count` independent solves (one per frame / camera stream of a batch) in ONE launch, one CTA each: pose7 [count][7]
 * in / out, visible[i] / proj_uv[i] per-problem arrays (the arrays or single entries may be NULL), n_inlier / n_iter
 * [count] (may be NULL).  SPFE_ERR_STATE if any problem hit the linearizeOplus throw (its n_iter is -1); the other
 * problems' results are still valid. */
int spfe_dust_pose_optimize_batch(spfe_ctx *ctx, const spfe_dust_pose *problems, int32_t count, double *pose7,
                                  uint8_t *const *visible, float *const *proj_uv, int32_t *n_inlier, int32_t *n_iter);
int spfe_dust_linearize(spfe_ctx *ctx, const spfe_dust_pose *p, const double *pose7, uint8_t *level, double *err,
                        float *proj_uv, double *J, double *Hb);

/* Changes the detection threshold (spfe_config.score_thresh, 0.007 upstream) for the batches submitted from now on.
 * The reference has no such knob; it exists for the optional global keypoint budget of a multi-GPU job: every rank
 * all-reduces a 64-bin score histogram (sp_orb_slam_b200/sharding.py, NCCL) and applies the common cut here. */
int spfe_set_score_threshold(spfe_ctx *ctx, float score_thresh);

/* SPFE_MATCH_PREV: forget the slot's previous frame (start of a new camera stream). */
int spfe_reset_stream(spfe_ctx *ctx, int32_t slot);

/* L2 distance between two 256-d descriptors (host, scalar). */
float spfe_l2(const float *a, const float *b);

/* ---- introspection used by tests / bench (not needed by the shim) ---- */
/* Copies a named intermediate device tensor of `slot` to dst (host).  Names:
 * "conv1a".."conv4b", "heads", "coarse" (fp16 NHWC), "score", "semi_dust",
 * "dense_dust", "heat_log" (f32), "argmax" (u8), "count" (i32 per frame).
 * Returns bytes copied or a negative error. */
int64_t spfe_debug_read(spfe_ctx *ctx, int32_t slot, const char *name, void *dst, size_t dst_bytes);
/* Device-side stopwatch on a slot's stream (CUDA events): start, enqueue any number of submits, stop.
 * spfe_timer_stop waits for the stream and returns the elapsed milliseconds between the two events. */
int spfe_timer_start(spfe_ctx *ctx, int32_t slot);
int spfe_timer_stop(spfe_ctx *ctx, int32_t slot, float *ms);
/* Parses a weight file with the library's own readers (no GPU needed).  Returns the number of
 * parameters (1300865 for SuperPoint) or a negative error; message in err[0..errcap). */
int64_t spfe_check_weights(const char *path, char *err, size_t errcap);
/* Parallel rounds the resolve phase of the last spfe_search_guided* call on this context ran (1 = no two queries
 * competed for a key point; tests). */
int32_t spfe_guided_last_rounds(const spfe_ctx *ctx);
/* Number of kernels this library has launched on the context so far. */
int64_t spfe_launch_count(const spfe_ctx *ctx);
/* CUDA events around the dominant kernel (the fused conv1a + conv1b) of every batch submitted from now on, the last 64
 * kept: spfe_dom_time returns their average launch duration -- the kernel timed INSIDE a long run (bench.py quotes it
 * against the sustained peak), where spfe_profile_device times it in an isolated 4-ms pass.  Default mode only. */
int spfe_dom_timing(spfe_ctx *ctx, int32_t enable);
int spfe_dom_time(spfe_ctx *ctx, float *avg_ms, int32_t *count);
/* Runs one device-resident batch with CUDA events around every stage; writes
 * up to `cap` (name, ms) pairs.  Returns the number of stages. */
typedef struct spfe_stage_time { char name[24]; float ms; double flop; double bytes; } spfe_stage_time;
int spfe_profile_device(spfe_ctx *ctx, int32_t slot, const void *d_gray, int32_t batch,
                        spfe_stage_time *stages, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SPFE_H_ */
