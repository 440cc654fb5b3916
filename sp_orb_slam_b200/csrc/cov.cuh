// computeCovariance on the device (reference orb_slam2/src/cv/sp_extractor.cpp:252-340), exact.
//
// The reference walks the keypoints in output order; for each it flood-fills (FIFO, 4-neighbours, strictly
// descending positive heat_inv) from the keypoint, marking pixels in a visited map that is SHARED by all keypoints
// of the frame (marked when popped, so a pixel can be queued -- and accumulated -- more than once), and forms
// score-weighted second moments of the visited pixels.  The shared map makes the result order dependent.
//
// Parallel restatement (every step is exact; floating-point operations use the reference's order and no fused
// multiply-add, so results are bit-identical to the C restatement in oracle/sp_post.c):
//   1. cov_flood_kernel: every keypoint k floods as if it were alone (its *lone* flood, a superset of its true flood)
//      and claims every popped pixel with atomicMin(owner[pixel], tag_0 | k).  One WARP per flood, expanding the FIFO
//      level by level (cov_warp_flood: exact pop order, one memory round trip per level instead of per pop), floods
//      handed out dynamically; queue (512 entries) and first-position map (64 x 64 window) of each flood live in shared
//      memory.  The rare flood that does not fit is redone with the big limits (2048 entries, 128 x 128 window); one
//      that does not fit even there sends its whole frame to the sequential path, which has no limits.
//   2. cov_resolve0_kernel (one warp per keypoint): k is *clean* iff every pixel of its lone flood is owned by k, i.e.
//      no lower-indexed lone flood touches it -- then the lone flood IS the reference's.  Clean keypoints compute their
//      moments (lanes load, ordered adds by shuffle) and stamp the true visited map; the others go to a pending list.
//   3. rounds r = 1..COV_ROUNDS of cov_claim_kernel + cov_resolve_kernel over the pending list: the keypoints still
//      pending re-claim their lone-flood pixels with tag_r (a later round's tag is smaller, so it overrides earlier
//      claims without clearing the map).  A pending keypoint that owns all its pixels has no unfinished lower-indexed
//      neighbour any more: its true flood is the flood constrained by the visited map (seeded into the private window
//      map on the pixels of the lone flood, the only ones it can reach), again by cov_warp_flood.
//      Stamps of keypoints finished in earlier rounds with a HIGHER index are harmless: they were clean while this
//      keypoint was pending, so their pixels are disjoint from this lone flood.
//   4. cov_replay_kernel: whatever is left (conflict chains deeper than COV_ROUNDS, overflowed frames) is replayed
//      sequentially in index order, one warp per frame, against the true visited map.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace spfe {

constexpr int COV_WIN = 128;     // big flood: window side around the keypoint (pixels)
constexpr int COV_QCAP = 2048;   // big flood: queue entries == stride of the per-keypoint pop lists in global memory
constexpr int COV_B_WARPS = 4;   // big floods (warps) per block
constexpr int COV_B_SMEM = COV_B_WARPS * (COV_WIN * COV_WIN * 2 + COV_QCAP * 2);  // 144 KB (two-byte first-position map)
constexpr int COV_S_WIN = 64;    // small flood (the common case)
constexpr int COV_S_QCAP = 1024;
constexpr int COV_S_WARPS = 32;  // small floods (warps) per block
constexpr int COV_S_SMEM = COV_S_WARPS * (COV_S_WIN * COV_S_WIN + COV_S_QCAP * 2);  // 192 KB (one-byte first-position map)
constexpr int COV_R_BIG_BLOCKS = 32;  // cov_resolve_kernel: blocks (of COV_B_WARPS active warps) that take the big floods
constexpr int COV_ROUNDS = 3;    // parallel conflict-resolution rounds before the sequential remainder
constexpr int COV_SEQ_QCAP = 32 * 1024;  // queue entries of the sequential path (128 KB of shared memory)
constexpr int COV_SEQ_BITMAP_WORDS = 16 * 1024;  // + 64 KB: the frame's visited bitmap if H*W <= 524288 pixels
// counters (ints) zeroed per batch: [0] flood work counter, [1] big-flood list length, [2] pending list length,
// [3] big-flood work counter
constexpr int COV_NCTR = 4;

struct CovArgs {
  const float *heat_inv;  // [B][H][W]
  const float *kp_xy;     // [B][cap][2]
  const int *count;       // [B]
  int *owner;             // [B][H*W]  claim of the lowest-indexed unfinished lone flood that popped the pixel (init 0x7F7F7F7F)
  uint32_t *visited;      // [B][vis_words] the reference's visited map as a bitmap
  int vis_words;          // (H*W + 31) / 32
  uint32_t *queue;        // [B][cap][COV_QCAP] pop list (pixel indices, duplicates included) of every LONE flood
  int *qlen;              // [B][cap]   its length
  int *done;              // [B][cap]   1 = response / cov2 / cov2_inv are final
  int *isbig;             // [B][cap]   1 = the lone flood needed the big limits
  int *ctr;               // [COV_NCTR]
  int *big;               // [B*cap]    keypoints (b*cap + k) whose lone flood needs the big limits
  int *pend;              // [B*cap]    keypoints not clean in round 0
  int *frame_flag;        // [B] 1 = a flood overflowed even the big limits: replay the whole frame sequentially
  float *response;        // [B][cap]
  float *cov2;            // [B][cap][2]
  float *cov2_inv;        // [B][cap][2]
  int *overflow;          // [1] frame index + 1 if even the sequential queue overflowed (reported as an error; cleared per batch)
  int *n_replay;          // [B][2] statistics: keypoints / pixels replayed sequentially
  int H, W, cap, B, round;
  int epoch_tag;          // (508 - batches since the owner map was cleared) << 22
  int force;              // test hook (SPFE_COV_FORCE): bit 0 = every small flood "does not fit", bit 1 = every big flood neither
};

// Claim value: smaller wins (atomicMin).  Later rounds of a batch override earlier ones, and later batches override
// earlier batches (epoch_tag falls from 508 << 22 to 0 over 509 batches), so the owner map is cleared only every 509
// batches instead of every batch; a stale claim of an older batch never equals a current one.
__device__ __forceinline__ int cov_tag(int epoch_tag, int round, int k) { return epoch_tag + ((63 - round) << 16) + k; }  // (63 << 16) < (1 << 22), k < 65536
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Moments of a popped-pixel list, in the reference's order of operations (sp_extractor.cpp:316-333); one thread.
__device__ __forceinline__ void cov_moments(const float *heat, const uint32_t *q, int n, int W, int cu, int cv, float *cov2,
                                            float *cov2_inv) {
  float sum = 0.0f;
  for (int i = 0; i < n; i++) sum = __fadd_rn(sum, __ldg(heat + q[i]));
  float sx = 0.0f, sy = 0.0f;
  for (int i = 0; i < n; i++) {
    const int pix = static_cast<int>(q[i]);
    const float du = __fsub_rn(static_cast<float>(pix % W), static_cast<float>(cu));
    const float dv = __fsub_rn(static_cast<float>(pix / W), static_cast<float>(cv));
    const float wgt = __fdiv_rn(__ldg(heat + pix), sum);
    sx = __fadd_rn(sx, __fmul_rn(wgt, __fmul_rn(du, du)));
    sy = __fadd_rn(sy, __fmul_rn(wgt, __fmul_rn(dv, dv)));
  }
  if (sx < 1.0f) sx = 1.0f;
  if (sy < 1.0f) sy = 1.0f;
  cov2[0] = sx;
  cov2[1] = sy;
  cov2_inv[0] = __fdiv_rn(1.0f, sx);
  cov2_inv[1] = __fdiv_rn(1.0f, sy);
}

// The same moments by a whole warp: lanes load and form the per-entry terms, the additions run in pop order in every
// lane (operands broadcast by shuffle).  get(i) -> (pixel index, u, v) of pop i.
template <class Get>
__device__ __forceinline__ void cov_moments_warp(const float *heat, int n, int W, int cu, int cv, int lane, float *cov2,
                                                 float *cov2_inv, Get get) {
  float sum = 0.0f;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int i = c0 + lane;
    float h = 0.0f;
    if (i < n) { int pix, u, v; get(i, pix, u, v); h = __ldg(heat + pix); }
    const int cn = min(32, n - c0);
    for (int j = 0; j < cn; j++) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, h, j));
  }
  float sx = 0.0f, sy = 0.0f;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int i = c0 + lane;
    float tx = 0.0f, ty = 0.0f;
    if (i < n) {
      int pix, u, v;
      get(i, pix, u, v);
      const float du = __fsub_rn(static_cast<float>(u), static_cast<float>(cu));
      const float dv = __fsub_rn(static_cast<float>(v), static_cast<float>(cv));
      const float wgt = __fdiv_rn(__ldg(heat + pix), sum);
      tx = __fmul_rn(wgt, __fmul_rn(du, du));
      ty = __fmul_rn(wgt, __fmul_rn(dv, dv));
    }
    const int cn = min(32, n - c0);
    for (int j = 0; j < cn; j++) {
      sx = __fadd_rn(sx, __shfl_sync(0xffffffffu, tx, j));
      sy = __fadd_rn(sy, __shfl_sync(0xffffffffu, ty, j));
    }
  }
  if (lane == 0) {
    if (sx < 1.0f) sx = 1.0f;
    if (sy < 1.0f) sy = 1.0f;
    cov2[0] = sx;
    cov2[1] = sy;
    cov2_inv[0] = __fdiv_rn(1.0f, sx);
    cov2_inv[1] = __fdiv_rn(1.0f, sy);
  }
}

// Exact warp-parallel FIFO flood.  The reference pops a queue in order and pushes the neighbours (left, up, right, down;
// boundary tests xx > 0, yy > 0, xx < w, yy < h) that are not yet visited (= popped) and strictly lower but positive.
// The queue order is reproduced level by level: the entries pushed by the pops of one level form the next level, in
// the same order; "p was visited when entry i popped" == "the first queue position of p is below i", and every such
// position is known before level i's expansion starts (entries created during the expansion lie behind the level).
// So per level the warp (1) records the first position of each of the level's pixels (fp, a window map private to the
// flood; 0xFFFF = never queued-and-popped, 0 = visited before the flood started), (2) expands 32 entries at a time --
// all heat_inv loads of a chunk are in flight together -- and appends the pushes with an ordered warp scan.
// A long flood costs one memory round trip per LEVEL (tens) instead of per pop (hundreds).
// Queue entry = wy * WIN + wx in a WIN x WIN window centred on the keypoint.  Returns the number of pops or -1 if the
// flood leaves the window or the queue.  fp must be initialised by the caller.
// Two encodings of the first-position map: absolute (uint16_t: position, 0xFFFF = unset, 0 = visited before the flood)
// for the big limits, and REL8 (uint8_t: 0 = unset, 1 = popped in an earlier level or visited before the flood,
// 2 + k = k-th entry of the level being expanded; a level of more than 253 entries does not fit) -- half the shared
// memory per flood, so twice as many floods in flight.
template <int WIN, bool REL8>
struct CovFp {
  using T = typename std::conditional<REL8, uint8_t, uint16_t>::type;
  static constexpr int BYTES = WIN * WIN * static_cast<int>(sizeof(T));
  static constexpr int UNSET = REL8 ? 0 : 0xFFFF;
  static constexpr int SEEN = REL8 ? 1 : 0;
};

template <int WIN, int QCAP, bool REL8>
__device__ __forceinline__ int cov_warp_flood(const float *heat, int W, int H, int ox, int oy, typename CovFp<WIN, REL8>::T *fp,
                                              uint16_t *sq, int lane, int *owner, int claim) {
  using FpT = typename CovFp<WIN, REL8>::T;
  if (lane == 0) sq[0] = static_cast<uint16_t>((WIN / 2) * WIN + WIN / 2);
  __syncwarp();
  int lo = 0, hi = 1;
  while (lo < hi) {
    if (REL8 && hi - lo > 253) return -1;
    for (int c0 = lo; c0 < hi; c0 += 32) {  // (1) first pop position of every pixel of this level
      const int i = c0 + lane;
      const bool act = i < hi;
      const int idx = act ? sq[i] : -1 - lane;
      const unsigned m = __match_any_sync(0xffffffffu, idx);
      if (act && (__ffs(m) - 1) == lane && fp[idx] == CovFp<WIN, REL8>::UNSET) fp[idx] = static_cast<FpT>(REL8 ? 2 + i - lo : i);
      __syncwarp();
    }
    int tail = hi;
    for (int c0 = lo; c0 < hi; c0 += 32) {  // (2) expand
      const int i = c0 + lane;
      int nidx[4];
      bool push[4] = {false, false, false, false};
      bool over = false;
      if (i < hi) {
        const int idx = sq[i];
        const int wx = idx & (WIN - 1), wy = idx / WIN;
        const int u = ox + wx, v = oy + wy, pix = v * W + u;
        if (owner != nullptr) atomicMin(owner + pix, claim);  // result unused: fire-and-forget reduction
        const bool ok[4] = {u - 1 > 0, v - 1 > 0, u + 1 < W, v + 1 < H};
        const int dpix[4] = {-1, -W, 1, W}, didx[4] = {-1, -WIN, 1, WIN};
        const bool inwin[4] = {wx > 0, wy > 0, wx < WIN - 1, wy < WIN - 1};
        const float here = __ldg(heat + pix);
        float hv[4];
#pragma unroll
        for (int d = 0; d < 4; d++) hv[d] = __ldg(heat + (ok[d] ? pix + dpix[d] : pix));
#pragma unroll
        for (int d = 0; d < 4; d++) {
          nidx[d] = idx + didx[d];
          if (!ok[d] || !(hv[d] > 0.0f && hv[d] < here)) continue;
          if (!inwin[d]) { over = true; continue; }
          const int f = fp[nidx[d]];
          push[d] = REL8 ? (f == 0 || (f >= 2 && f - 2 + lo > i)) : (f > i);
        }
      }
      const int cnt = static_cast<int>(push[0]) + push[1] + push[2] + push[3];
      int off = cnt;  // inclusive scan over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, off, o);
        if (lane >= o) off += t;
      }
      const int total = __shfl_sync(0xffffffffu, off, 31);
      if (__any_sync(0xffffffffu, over) || tail + total > QCAP) return -1;
      int w = tail + off - cnt;
#pragma unroll
      for (int d = 0; d < 4; d++)
        if (push[d]) sq[w++] = static_cast<uint16_t>(nidx[d]);
      tail += total;
      __syncwarp();
    }
    if (REL8) {  // the level is done: its pixels are "popped in an earlier level" from now on
      for (int i = lo + lane; i < hi; i += 32) fp[sq[i]] = 1;
      __syncwarp();
    }
    lo = hi;
    hi = tail;
  }
  return hi;
}

template <int WIN, bool REL8>
__device__ __forceinline__ void cov_fp_init(typename CovFp<WIN, REL8>::T *fp, int lane) {
  uint4 *p = reinterpret_cast<uint4 *>(fp);
  const unsigned v = REL8 ? 0u : ~0u;
  for (int i = lane; i < CovFp<WIN, REL8>::BYTES / 16; i += 32) p[i] = make_uint4(v, v, v, v);
  __syncwarp();
}

// Phase 1: lone flood of every keypoint, one warp per flood, floods handed out dynamically.  BIG = false: all keypoints,
// small limits, the ones that do not fit are listed in a.big; BIG = true: that list with the big limits.
template <int WIN, int QCAP, int WARPS, bool BIG>
__global__ void __launch_bounds__(WARPS * 32) cov_flood_kernel(const CovArgs a) {
  extern __shared__ uint8_t cov_smem[];
  constexpr bool REL8 = !BIG;
  using Fp = CovFp<WIN, REL8>;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  typename Fp::T *fp = reinterpret_cast<typename Fp::T *>(cov_smem + static_cast<size_t>(w) * (Fp::BYTES + QCAP * 2));
  uint16_t *sq = reinterpret_cast<uint16_t *>(cov_smem + static_cast<size_t>(w) * (Fp::BYTES + QCAP * 2) + Fp::BYTES);
  const int W = a.W, H = a.H, total = BIG ? a.ctr[1] : a.B * a.cap;
  const size_t px = static_cast<size_t>(H) * W;
  for (;;) {
    int e = 0;
    if (lane == 0) e = atomicAdd(a.ctr + (BIG ? 3 : 0), 1);
    e = __shfl_sync(0xffffffffu, e, 0);
    if (e >= total) break;
    const int ki = BIG ? a.big[e] : e;
    const int b = ki / a.cap, k = ki - b * a.cap;
    if (k >= a.count[b]) continue;
    const float *xy = a.kp_xy + static_cast<size_t>(ki) * 2;
    const int ox = static_cast<int>(xy[0]) - WIN / 2, oy = static_cast<int>(xy[1]) - WIN / 2;
    cov_fp_init<WIN, REL8>(fp, lane);
    int n = cov_warp_flood<WIN, QCAP, REL8>(a.heat_inv + b * px, W, H, ox, oy, fp, sq, lane, a.owner + b * px, cov_tag(a.epoch_tag, 0, k));
    if (a.force & (BIG ? 2 : 1)) n = -1;
    __syncwarp();
    if (n < 0) {
      if (lane == 0) {
        a.qlen[ki] = -1;
        if (BIG) a.frame_flag[b] = 1;
        else a.big[atomicAdd(a.ctr + 1, 1)] = ki;
      }
      continue;
    }
    uint32_t *q = a.queue + static_cast<size_t>(ki) * COV_QCAP;
    for (int i = lane; i < n; i += 32) {
      const int idx = sq[i];
      q[i] = static_cast<uint32_t>((oy + idx / WIN) * W + ox + (idx & (WIN - 1)));
    }
    if (lane == 0) {
      a.qlen[ki] = n;
      a.isbig[ki] = BIG;
    }
    __syncwarp();
  }
}

// Phase 2: one warp per keypoint.  Clean keypoints are final; the others are queued for the rounds.
__global__ void __launch_bounds__(256) cov_resolve0_kernel(const CovArgs a) {
  const int lane = threadIdx.x & 31;
  const int ki = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ki >= a.B * a.cap) return;
  const int b = ki / a.cap, k = ki - b * a.cap;
  if (k >= a.count[b]) return;
  const size_t px = static_cast<size_t>(a.H) * a.W;
  const float *heat = a.heat_inv + b * px;
  const int *owner = a.owner + b * px;
  const uint32_t *q = a.queue + static_cast<size_t>(ki) * COV_QCAP;
  const float *xy = a.kp_xy + static_cast<size_t>(ki) * 2;
  const int cu = static_cast<int>(xy[0]), cv = static_cast<int>(xy[1]), W = a.W;
  if (lane == 0) {
    a.response[ki] = heat[cv * W + cu];  // kp.response = heat_inv(v, u), sp_extractor.cpp:271
    a.done[ki] = 0;
  }
  if (a.frame_flag[b]) return;  // whole frame goes through the sequential path
  const int n = a.qlen[ki], mine = cov_tag(a.epoch_tag, 0, k);
  bool dirty = false;
  for (int i = lane; i < n; i += 32) dirty |= (owner[q[i]] != mine);
  if (__any_sync(0xffffffffu, dirty)) {
    if (lane == 0) a.pend[atomicAdd(a.ctr + 2, 1)] = ki;
    return;
  }
  cov_moments_warp(heat, n, W, cu, cv, lane, a.cov2 + static_cast<size_t>(ki) * 2, a.cov2_inv + static_cast<size_t>(ki) * 2,
                   [&](int i, int &pix, int &u, int &v) { pix = static_cast<int>(q[i]); v = pix / W; u = pix - v * W; });
  uint32_t *visited = a.visited + static_cast<size_t>(b) * a.vis_words;
  for (int i = lane; i < n; i += 32) atomicOr(visited + (q[i] >> 5), 1u << (q[i] & 31));
  if (lane == 0) a.done[ki] = 1;
}

// Round r, step 1: the keypoints still pending re-claim the pixels of their lone floods (one warp per keypoint).
__global__ void __launch_bounds__(256) cov_claim_kernel(const CovArgs a) {
  const int lane = threadIdx.x & 31, n_pend = a.ctr[2];
  for (int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n_pend; e += gridDim.x * (blockDim.x >> 5)) {
    const int ki = a.pend[e];
    if (a.done[ki]) continue;
    const int b = ki / a.cap, k = ki - b * a.cap;
    if (a.frame_flag[b]) continue;
    int *owner = a.owner + static_cast<size_t>(b) * a.H * a.W;
    const uint32_t *q = a.queue + static_cast<size_t>(ki) * COV_QCAP;
    const int n = a.qlen[ki], mine = cov_tag(a.epoch_tag, a.round, k);
    for (int i = lane; i < n; i += 32) atomicMin(owner + q[i], mine);
  }
}

// Round r, step 2: a pending keypoint that owns all its pixels has no unfinished lower-indexed neighbour: flood it
// against the visited map (seeded into fp on the pixels of the lone flood, the only ones it can reach).
template <int WIN, int QCAP, bool REL8>
__device__ __forceinline__ void cov_resolve_warp(const CovArgs &a, typename CovFp<WIN, REL8>::T *fp, uint16_t *sq, int lane, int first,
                                                 int stride, bool big) {
  const int W = a.W, H = a.H, n_pend = a.ctr[2];
  for (int e = first; e < n_pend; e += stride) {
    const int ki = a.pend[e];
    if (a.done[ki]) continue;
    const int b = ki / a.cap, k = ki - b * a.cap;
    if (a.frame_flag[b]) continue;
    if ((a.isbig[ki] != 0) != big) continue;  // the blocks of the other kind handle it
    const size_t px = static_cast<size_t>(H) * W;
    const float *heat = a.heat_inv + b * px;
    const int *owner = a.owner + b * px;
    uint32_t *visited = a.visited + static_cast<size_t>(b) * a.vis_words;
    const uint32_t *q = a.queue + static_cast<size_t>(ki) * COV_QCAP;
    const int n = a.qlen[ki], mine = cov_tag(a.epoch_tag, a.round, k);
    bool blocked = false;
    for (int i = lane; i < n; i += 32) blocked |= (owner[q[i]] != mine);
    if (__any_sync(0xffffffffu, blocked)) continue;
    const float *xy = a.kp_xy + static_cast<size_t>(ki) * 2;
    const int cu = static_cast<int>(xy[0]), cv = static_cast<int>(xy[1]);
    const int ox = cu - WIN / 2, oy = cv - WIN / 2;
    cov_fp_init<WIN, REL8>(fp, lane);
    for (int i = lane; i < n; i += 32) {
      const int pix = static_cast<int>(q[i]), v = pix / W, u = pix - v * W;
      if ((__ldcg(visited + (pix >> 5)) >> (pix & 31)) & 1u) fp[(v - oy) * WIN + (u - ox)] = CovFp<WIN, REL8>::SEEN;
    }
    __syncwarp();
    const int tail = cov_warp_flood<WIN, QCAP, REL8>(heat, W, H, ox, oy, fp, sq, lane, nullptr, 0);
    __syncwarp();
    if (tail <= 0) continue;  // does not fit: left to the sequential remainder
    cov_moments_warp(heat, tail, W, cu, cv, lane, a.cov2 + static_cast<size_t>(ki) * 2, a.cov2_inv + static_cast<size_t>(ki) * 2,
                     [&](int i, int &pix, int &u, int &v) {
                       const int idx = sq[i];
                       u = ox + (idx & (WIN - 1));
                       v = oy + idx / WIN;
                       pix = v * W + u;
                     });
    for (int i = lane; i < tail; i += 32) {
      const int idx = sq[i];
      const int pix = (oy + idx / WIN) * W + ox + (idx & (WIN - 1));
      atomicOr(visited + (pix >> 5), 1u << (pix & 31));
    }
    if (lane == 0) a.done[ki] = 1;
    __syncwarp();
  }
}

// One launch per round: the last COV_R_BIG_BLOCKS blocks take the keypoints whose lone flood needed the big limits
// (COV_B_WARPS warps of the block work, with the big window / queue), all other blocks the common small ones.
__global__ void __launch_bounds__(COV_S_WARPS * 32) cov_resolve_kernel(const CovArgs a) {
  extern __shared__ uint8_t cov_smem[];
  static_assert(COV_B_SMEM <= COV_S_SMEM, "both kinds of block use the same launch configuration");
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_small = static_cast<int>(gridDim.x) - COV_R_BIG_BLOCKS;
  if (static_cast<int>(blockIdx.x) < n_small) {
    using Fp = CovFp<COV_S_WIN, true>;
    uint8_t *base = cov_smem + static_cast<size_t>(w) * (Fp::BYTES + COV_S_QCAP * 2);
    cov_resolve_warp<COV_S_WIN, COV_S_QCAP, true>(a, base, reinterpret_cast<uint16_t *>(base + Fp::BYTES), lane, blockIdx.x * COV_S_WARPS + w,
                                                  n_small * COV_S_WARPS, false);
  } else if (w < COV_B_WARPS && a.ctr[1] > 0) {
    using Fp = CovFp<COV_WIN, false>;
    uint8_t *base = cov_smem + static_cast<size_t>(w) * (Fp::BYTES + COV_QCAP * 2);
    cov_resolve_warp<COV_WIN, COV_QCAP, false>(a, reinterpret_cast<uint16_t *>(base), reinterpret_cast<uint16_t *>(base + Fp::BYTES), lane,
                                               (blockIdx.x - n_small) * COV_B_WARPS + w, COV_R_BIG_BLOCKS * COV_B_WARPS, true);
  }
}

// Phase 4: sequential replay (one warp per frame) of the keypoints that are still not final, in index order, against
// the true visited map.  A pop is a chain of dependent accesses, so everything it touches is kept close: the queue
// (entries packed v << 16 | u: no divisions) and, for frames up to 524288 pixels, the frame's visited bitmap in
// shared memory; heat_inv on the read-only path.  Lane 0 runs the floods; all lanes gather the list (ordered ballot
// compaction) and compute the per-entry moment terms, lane 0 only performs the ordered additions.
__global__ void __launch_bounds__(32) cov_replay_kernel(const CovArgs a) {
  extern __shared__ uint32_t cov_rsmem[];
  uint32_t *rq = cov_rsmem;                    // [COV_SEQ_QCAP]
  uint32_t *s_vis = cov_rsmem + COV_SEQ_QCAP;  // [COV_SEQ_BITMAP_WORDS]
  __shared__ uint16_t s_dirty[4096];
  __shared__ float s_t0[1024], s_t1[1024];
  const int b = blockIdx.x, lane = threadIdx.x;
  uint32_t *g_vis = a.visited + static_cast<size_t>(b) * a.vis_words;
  const size_t px = static_cast<size_t>(a.H) * a.W;
  const float *heat = a.heat_inv + b * px;
  const int n_kp = a.count[b], W = a.W, H = a.H;
  // ordered list of the keypoints to replay, gathered in windows of at most 4096 (any key-point budget works)
  int k_next = 0, n_total = 0, stat_p = 0;
  bool vis_loaded = false;
  const bool in_smem = a.vis_words <= COV_SEQ_BITMAP_WORDS;
  uint32_t *vis = in_smem ? s_vis : g_vis;
  if (lane == 0) {
    a.n_replay[2 * b] = 0;
    a.n_replay[2 * b + 1] = 0;
  }
  while (k_next < n_kp) {
  int n_dirty = 0;
  for (; k_next < n_kp && n_dirty <= 4096 - 32; k_next += 32) {
    const int k = k_next + lane;
    const bool d = k < n_kp && !a.done[static_cast<size_t>(b) * a.cap + k];
    const unsigned m = __ballot_sync(0xffffffffu, d);
    if (d) s_dirty[n_dirty + __popc(m & ((1u << lane) - 1))] = static_cast<uint16_t>(k);
    n_dirty += __popc(m);
  }
  __syncwarp();
  if (n_dirty == 0) continue;
  n_total += n_dirty;
  if (!vis_loaded) {
    if (in_smem)
      for (int i = lane; i < a.vis_words; i += 32) s_vis[i] = g_vis[i];
    vis_loaded = true;
    __syncwarp();
  }
  for (int di = 0; di < n_dirty; di++) {
    const int k = s_dirty[di];
    const size_t ki = static_cast<size_t>(b) * a.cap + k;
    const float *xy = a.kp_xy + ki * 2;
    const int cu = static_cast<int>(xy[0]), cv = static_cast<int>(xy[1]);
    {  // all lanes warm L1 with a 32-row x 96-pixel neighbourhood of heat_inv before lane 0 starts the dependent chain
      const int v = min(max(cv - 16 + lane, 0), H - 1);
      prefetch_l1(heat + v * W + max(cu - 32, 0));
      prefetch_l1(heat + v * W + cu);
      prefetch_l1(heat + v * W + min(cu + 32, W - 1));
    }
    int tail = 0;
    if (lane == 0) {
      int head = 0;
      rq[tail++] = (cv << 16) | cu;
      while (head < tail) {
        const uint32_t e = rq[head++];
        const int u = e & 0xFFFF, v = e >> 16, pix = v * W + u;
        vis[pix >> 5] |= 1u << (pix & 31);
        const bool ok[4] = {u - 1 > 0, v - 1 > 0, u + 1 < W, v + 1 < H};  // reference order: left, up, right, down
        const int cand[4] = {ok[0] ? pix - 1 : pix, ok[1] ? pix - W : pix, ok[2] ? pix + 1 : pix, ok[3] ? pix + W : pix};
        const uint32_t ce[4] = {e - 1, e - 0x10000u, e + 1, e + 0x10000u};
        const float here = __ldg(heat + pix);
        float hv[4];
#pragma unroll
        for (int d = 0; d < 4; d++) hv[d] = __ldg(heat + cand[d]);
#pragma unroll
        for (int d = 0; d < 4; d++) {
          if (!ok[d] || !(hv[d] > 0.0f && hv[d] < here)) continue;
          if ((vis[cand[d] >> 5] >> (cand[d] & 31)) & 1u) continue;
          if (tail >= COV_SEQ_QCAP) { tail = -1; break; }
          rq[tail++] = ce[d];
        }
        if (tail < 0) break;
      }
    }
    tail = __shfl_sync(0xffffffffu, tail, 0);
    if (tail < 0) {
      if (lane == 0) atomicMax(a.overflow, b + 1);  // which frame: reported by spfe_wait
      continue;
    }
    // moments (sp_extractor.cpp:316-333): sum of scores, then sum of (score / sum) * delta^2, both in pop order
    float sum = 0.0f;
    for (int c0 = 0; c0 < tail; c0 += 1024) {
      const int cn = min(1024, tail - c0);
      for (int i = lane; i < cn; i += 32) { const uint32_t e = rq[c0 + i]; s_t0[i] = __ldg(heat + (e >> 16) * W + (e & 0xFFFF)); }
      __syncwarp();
      if (lane == 0) for (int i = 0; i < cn; i++) sum = __fadd_rn(sum, s_t0[i]);
      __syncwarp();
    }
    sum = __shfl_sync(0xffffffffu, sum, 0);
    float sx = 0.0f, sy = 0.0f;
    for (int c0 = 0; c0 < tail; c0 += 1024) {
      const int cn = min(1024, tail - c0);
      for (int i = lane; i < cn; i += 32) {
        const uint32_t e = rq[c0 + i];
        const int u = e & 0xFFFF, v = e >> 16;
        const float du = __fsub_rn(static_cast<float>(u), static_cast<float>(cu));
        const float dv = __fsub_rn(static_cast<float>(v), static_cast<float>(cv));
        const float wgt = __fdiv_rn(__ldg(heat + v * W + u), sum);
        s_t0[i] = __fmul_rn(wgt, __fmul_rn(du, du));
        s_t1[i] = __fmul_rn(wgt, __fmul_rn(dv, dv));
      }
      __syncwarp();
      if (lane == 0)
        for (int i = 0; i < cn; i++) { sx = __fadd_rn(sx, s_t0[i]); sy = __fadd_rn(sy, s_t1[i]); }
      __syncwarp();
    }
    if (lane == 0) {
      if (sx < 1.0f) sx = 1.0f;
      if (sy < 1.0f) sy = 1.0f;
      a.cov2[ki * 2] = sx;
      a.cov2[ki * 2 + 1] = sy;
      a.cov2_inv[ki * 2] = __fdiv_rn(1.0f, sx);
      a.cov2_inv[ki * 2 + 1] = __fdiv_rn(1.0f, sy);
      a.done[ki] = 1;
    }
    stat_p += tail;
  }
  }  // windows of the dirty list
  if (lane == 0 && n_total > 0) {
    a.n_replay[2 * b] = n_total;
    a.n_replay[2 * b + 1] = stat_p;
  }
}

}  // namespace spfe
