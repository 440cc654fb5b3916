// Guided (cell-grid) searches on the device -- SURVEY.md section 8(f) rank 2.
//
// Replaces the greedy candidate loops of
//   SPMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, th_dist)   sp_matcher.cpp:344-432
//   SPMatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono)        sp_matcher.cpp:1439-1543
//   the dust-track patch association                                               tracker_dust.cpp:112-172
// which all have the same shape: for every map point, in order, look up the keypoints of a few occ_grid cells around
// its projection (Frame::GetFeaturesInArea, frame.cpp:382-420, or the 2 x 2 cells at floor(proj)), skip the keypoints
// that already carry an observed map point, take the nearest descriptor (first one on ties), accept it under a
// distance rule, and make it unavailable to the map points that follow.
//
// ONE launch (guided_kernel), two phases:
//   candidates   one warp per map point: candidate keypoints in the reference's order (ix outer, iy inner, ordered
//                ballot compaction) and their 256-d L2 distances (lanes own 8 dimensions each).
//   resolve      the order-dependent part, exact, run by the LAST CTA to finish the first phase (atomic ticket): the
//                greedy loop is the fixpoint of "a map point decides once it is the lowest-indexed undecided map point
//                on every still-available candidate of its" (claims by atomicMin with a per-round tag, as in cov.cuh);
//                map points deciding in the same round have disjoint available candidates.  Chains deeper than
//                GUIDED_ROUNDS are finished sequentially by one thread.  The results are then written straight into
//                the caller's page-locked result block (mapped host memory): no device -> host copy is enqueued.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spfe {

constexpr int GUIDED_CAND = 64;    // candidate keypoints per map point (cells of a 2r+1 window, r <= ~24 px)
constexpr int GUIDED_ROUNDS = 256;  // <= 2038 (claim tags, guided_resolve_kernel)

struct GuidedArgs {
  int mode, m, n, grid_rows, grid_cols;
  const float *qdesc, *qxy, *qr;
  const uint8_t *qvalid, *qblocks;  // may be null
  const float *kdesc, *kp_un;
  const int16_t *occ;
  const uint8_t *taken_in;  // [n] keypoints that already carry an observed map point
  uint8_t *taken;   // [n] out: taken_in plus the keypoints claimed by this search (initialised by guided_cand_kernel)
  int *kpmin;       // [n] scratch, 0x7F7F7F7F
  int *cand;        // [m][GUIDED_CAND]
  float *cdist;     // [m][GUIDED_CAND]
  int *ncand;       // [m]
  uint8_t *decided; // [m]
  int *q2kp;        // [m]
  float *qdist;     // [m]
  int *overflow;    // [1] a map point had more than GUIDED_CAND candidates
  int *ticket;      // [1] CTAs that finished the candidate phase (0 at launch)
  int *h_q2kp;      // mapped host memory: results of the call (the last CTA copies them out)
  float *h_qdist;
  uint8_t *h_taken;
  int *h_overflow;  // [2] overflow flag, parallel rounds the resolve phase ran
  int smem_state;   // 1: the launch carries n * 5 + 16 bytes of dynamic shared memory for the resolve phase's taken / claim arrays
  float min_x, min_y, best_init, th_le, th_lt, c2;
};

__device__ __forceinline__ void guided_candidates(const GuidedArgs &a, int i, int lane) {
  if (lane == 0) {
    a.q2kp[i] = -1;
    a.qdist[i] = 0.0f;
  }
  if (a.qvalid != nullptr && !a.qvalid[i]) {
    if (lane == 0) { a.ncand[i] = 0; a.decided[i] = 1; }
    return;
  }
  const float x = a.qxy[2 * i], y = a.qxy[2 * i + 1];
  int x0, x1, y0, y1;
  float r = 0.0f;
  bool none = false;
  if (a.mode == 0) {  // Frame::GetFeaturesInArea, frame.cpp:387-401
    r = a.qr[i];
    x0 = max(0, static_cast<int>(floorf(__fdiv_rn(x - a.min_x - r, 8.0f))));
    x1 = min(a.grid_cols - 1, static_cast<int>(ceilf(__fdiv_rn(x - a.min_x + r, 8.0f))));
    y0 = max(0, static_cast<int>(floorf(__fdiv_rn(y - a.min_y - r, 8.0f))));
    y1 = min(a.grid_rows - 1, static_cast<int>(ceilf(__fdiv_rn(y - a.min_y + r, 8.0f))));
    none = x0 >= a.grid_cols || x1 < 0 || y0 >= a.grid_rows || y1 < 0;
  } else {            // tracker_dust.cpp:118-127: cells (u + du, v + dv), du outer
    x0 = static_cast<int>(floorf(x)); x1 = x0 + 1;
    y0 = static_cast<int>(floorf(y)); y1 = y0 + 1;
  }
  const int ny = y1 - y0 + 1, ncells = none ? 0 : max(0, (x1 - x0 + 1)) * max(0, ny);
  int cnt = 0;
  for (int c0 = 0; c0 < ncells; c0 += 32) {
    const int c = c0 + lane;
    bool ok = false;
    int idx = -1;
    if (c < ncells) {
      const int ix = x0 + c / ny, iy = y0 + c % ny;
      if (ix >= 0 && ix < a.grid_cols && iy >= 0 && iy < a.grid_rows) {
        idx = a.occ[iy * a.grid_cols + ix];
        if (idx >= 0 && idx < a.n) ok = a.mode != 0 || (fabsf(a.kp_un[2 * idx] - x) < r && fabsf(a.kp_un[2 * idx + 1] - y) < r);
      }
    }
    const unsigned msk = __ballot_sync(0xffffffffu, ok);
    const int pos = cnt + __popc(msk & ((1u << lane) - 1));
    if (ok && pos < GUIDED_CAND) a.cand[static_cast<size_t>(i) * GUIDED_CAND + pos] = idx;
    cnt += __popc(msk);
  }
  if (cnt > GUIDED_CAND) {
    if (lane == 0) atomicExch(a.overflow, 1);
    cnt = GUIDED_CAND;
  }
  __syncwarp();
  // SPMatcher::DescriptorDistance = cv::norm(a, b, NORM_L2) (sp_matcher.cpp:1636-1640)
  const float4 *q4 = reinterpret_cast<const float4 *>(a.qdesc + static_cast<size_t>(i) * 256) + lane * 2;
  const float4 qa = q4[0], qb = q4[1];
  for (int c = 0; c < cnt; c++) {
    const int idx = a.cand[static_cast<size_t>(i) * GUIDED_CAND + c];
    const float4 *k4 = reinterpret_cast<const float4 *>(a.kdesc + static_cast<size_t>(idx) * 256) + lane * 2;
    const float4 ka = __ldg(k4), kb = __ldg(k4 + 1);
    float s = 0.0f, d;
    d = qa.x - ka.x; s = fmaf(d, d, s);
    d = qa.y - ka.y; s = fmaf(d, d, s);
    d = qa.z - ka.z; s = fmaf(d, d, s);
    d = qa.w - ka.w; s = fmaf(d, d, s);
    d = qb.x - kb.x; s = fmaf(d, d, s);
    d = qb.y - kb.y; s = fmaf(d, d, s);
    d = qb.z - kb.z; s = fmaf(d, d, s);
    d = qb.w - kb.w; s = fmaf(d, d, s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) a.cdist[static_cast<size_t>(i) * GUIDED_CAND + c] = sqrtf(s);
  }
  if (lane == 0) { a.ncand[i] = cnt; a.decided[i] = 0; }
}

// The reference's per-map-point decision, given which keypoints are unavailable (sp_matcher.cpp:380-428,
// :1499-1532, tracker_dust.cpp:117-168).
__device__ __forceinline__ void guided_decide(const GuidedArgs &a, int i, volatile uint8_t *taken) {
  const int nc = a.ncand[i];
  float best = a.best_init;
  int bi = -1;
  for (int c = 0; c < nc; c++) {
    const int kp = a.cand[static_cast<size_t>(i) * GUIDED_CAND + c];
    if (taken[kp]) continue;
    const float d = a.cdist[static_cast<size_t>(i) * GUIDED_CAND + c];
    if (d < best) { best = d; bi = kp; }
  }
  a.decided[i] = 1;
  if (bi < 0) return;
  bool accept = best <= a.th_le;
  if (!accept) {
    float thr = a.th_lt;
    if (a.c2 > 0.0f) {  // tracking::map::match_adaptive, sp_matcher.cpp:416-423
      const float du = __fsub_rn(a.kp_un[2 * bi], a.qxy[2 * i]), dv = __fsub_rn(a.kp_un[2 * bi + 1], a.qxy[2 * i + 1]);
      const float duv = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
      thr = __fdiv_rn(__fmul_rn(1.2f, a.c2), __fadd_rn(a.c2, duv));
    }
    accept = best < thr;
  }
  if (!accept) return;
  a.q2kp[i] = bi;
  a.qdist[i] = best;
  if (a.qblocks == nullptr || a.qblocks[i]) taken[bi] = 1;
}

__device__ __forceinline__ void guided_resolve(const GuidedArgs &a) {
  extern __shared__ __align__(16) uint8_t guided_smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  // The two arrays every round hammers -- which key points are taken, who claims which -- live in shared memory when
  // they fit (n * 5 bytes): a round then costs shared-memory latencies instead of L2 round trips (the kernel is a chain
  // of dependent rounds on one SM: 33 -> ~15 us at 1000 map points).
  uint8_t *g_taken = a.taken;
  int *g_kpmin = a.kpmin;
  if (a.smem_state) {
    g_taken = guided_smem;
    g_kpmin = reinterpret_cast<int *>(guided_smem + ((a.n + 15) & ~15));
    for (int k = tid; k < a.n; k += nt) { g_taken[k] = a.taken[k]; g_kpmin[k] = 0x7F7F7F7F; }
    __syncthreads();
  }
  volatile uint8_t *taken = g_taken;
  volatile int *kpmin = g_kpmin;
  volatile uint8_t *decided = a.decided;
  bool left = true;
  int rounds = 0;
  for (int round = 0; round < GUIDED_ROUNDS && left; round++, rounds++) {
    const int tag = (2038 - round) << 20;  // every tag stays below the cleared value 0x7F7F7F7F, so round 0 already claims
    for (int i = tid; i < a.m; i += nt) {
      if (decided[i]) continue;
      const int nc = a.ncand[i];
      for (int c = 0; c < nc; c++) {
        const int kp = a.cand[static_cast<size_t>(i) * GUIDED_CAND + c];
        if (!taken[kp]) atomicMin(g_kpmin + kp, tag | i);
      }
    }
    __syncthreads();
    int und = 0;
    for (int i = tid; i < a.m; i += nt) {
      if (decided[i]) continue;
      const int nc = a.ncand[i];
      bool ready = true;
      for (int c = 0; c < nc; c++) {
        const int kp = a.cand[static_cast<size_t>(i) * GUIDED_CAND + c];
        if (!taken[kp] && kpmin[kp] != (tag | i)) ready = false;
      }
      if (ready) guided_decide(a, i, taken);
      else und = 1;
    }
    left = __syncthreads_or(und) != 0;
  }
  if (left && tid == 0)  // conflict chains deeper than GUIDED_ROUNDS: the plain sequential loop for the rest
    for (int i = 0; i < a.m; i++)
      if (!decided[i]) guided_decide(a, i, taken);
  if (a.smem_state) {
    __syncthreads();
    for (int k = tid; k < a.n; k += nt) a.taken[k] = g_taken[k];
  }
  if (tid == 0) a.h_overflow[1] = rounds;
}

constexpr int GUIDED_THREADS = 1024;  // 32 map points per CTA in the candidate phase; the resolve phase uses all of them

__global__ void __launch_bounds__(GUIDED_THREADS) guided_kernel(const GuidedArgs a) {
  __shared__ int s_last;
  for (int k = blockIdx.x * GUIDED_THREADS + threadIdx.x; k < a.n; k += gridDim.x * GUIDED_THREADS) a.taken[k] = a.taken_in[k];
  const int lane = threadIdx.x & 31, i = blockIdx.x * (GUIDED_THREADS / 32) + (threadIdx.x >> 5);
  if (i < a.m) guided_candidates(a, i, lane);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1) == static_cast<int>(gridDim.x) - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();  // every other CTA's candidates / taken copy are visible now
  guided_resolve(a);
  __syncthreads();
  for (int k = threadIdx.x; k < a.m; k += GUIDED_THREADS) {
    a.h_q2kp[k] = a.q2kp[k];
    a.h_qdist[k] = a.qdist[k];
  }
  for (int k = threadIdx.x; k < a.n; k += GUIDED_THREADS) a.h_taken[k] = a.taken[k];
  if (threadIdx.x == 0) *a.h_overflow = *a.overflow;
  __threadfence_system();
}

}  // namespace spfe
