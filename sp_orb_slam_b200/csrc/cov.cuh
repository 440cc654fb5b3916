// computeCovariance on the device (reference orb_slam2/src/cv/sp_extractor.cpp:252-340), exact.
//
// The reference walks the keypoints in output order; for each it flood-fills (FIFO, 4-neighbours, strictly
// descending positive heat_inv) from the keypoint, marking pixels in a visited map that is SHARED by all keypoints
// of the frame (marked when popped, so a pixel can be queued -- and accumulated -- more than once), and forms
// score-weighted second moments of the visited pixels.  The shared map makes the result order dependent.
//
// Parallel restatement:
//   1. cov_flood_kernel: every keypoint k floods in parallel as if it were alone (private visited bitmap) and
//      claims every popped pixel with red.min(owner[pixel], k).
//   2. cov_finish_kernel: keypoint k is *clean* iff every pixel of its flood ends up owned by k, i.e. no
//      lower-indexed keypoint's flood (a superset of its true, constrained flood) touches it -- then the lone flood
//      IS the reference's.  Clean keypoints compute their moments in pop order and stamp the true visited map.
//   3. cov_replay_kernel: the remaining (few per cent) keypoints are replayed sequentially in index order, one thread
//      per frame, against the true visited map.  Stamping the clean floods of LATER keypoints beforehand is
//      harmless: being clean means no earlier keypoint can reach them.
// A flood is a chain of dependent loads, so its state lives where latency is lowest: queue and visited bitmap of
// each flood in shared memory (a 128x128-pixel window around the keypoint), heat_inv through the read-only path,
// owner claims as fire-and-forget reductions.  A flood that leaves its window or queue (never seen: floods are
// tens of pixels) sends the whole frame to the sequential path, which has no such limits.
// Floating-point operations use the reference's order and no fused multiply-add: results are bit-identical to the
// C restatement in oracle/sp_post.c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spfe {

constexpr int COV_WIN = 128;     // flood window side around the keypoint (pixels)
constexpr int COV_QCAP = 2048;   // queue entries of a windowed flood
constexpr int COV_TPB = 32;      // floods (threads) per block in cov_flood_kernel
constexpr int COV_SMEM = COV_TPB * (COV_QCAP * 2 + COV_WIN * COV_WIN / 8);  // 192 KB
constexpr int COV_SEQ_QCAP = 32 * 1024;  // queue entries of the sequential path (128 KB of shared memory)
constexpr int COV_SEQ_BITMAP_WORDS = 16 * 1024;  // + 64 KB: the frame's visited bitmap if H*W <= 524288 pixels

struct CovArgs {
  const float *heat_inv;  // [B][H][W]
  const float *kp_xy;     // [B][cap][2]
  const int *count;       // [B]
  int *owner;             // [B][H*W]  lowest keypoint index that popped the pixel (init 0x7F7F7F7F)
  uint32_t *visited;      // [B][vis_words] the reference's visited map as a bitmap, built by kernels 2 and 3
  int vis_words;          // (H*W + 31) / 32
  uint32_t *queue;        // [B][cap][COV_QCAP] popped pixels of every flood, in order (duplicates included)
  int *qlen;              // [B][cap]   entries used; -1 = replay
  int *frame_flag;        // [B] 1 = a windowed flood overflowed: replay the whole frame sequentially
  float *response;        // [B][cap]
  float *cov2;            // [B][cap][2]
  float *cov2_inv;        // [B][cap][2]
  int *overflow;          // [1] set if even the sequential queue overflowed (reported as an error)
  int *n_replay;          // [B][2] statistics: keypoints / pixels replayed sequentially
  int H, W, cap;
};

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Moments of a popped-pixel list, in the reference's order of operations (sp_extractor.cpp:316-333).
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

// Phase 1: lone flood of every keypoint; one thread per keypoint, state in shared memory.
__global__ void __launch_bounds__(COV_TPB) cov_flood_kernel(const CovArgs a) {
  extern __shared__ uint8_t cov_smem[];
  uint16_t *sq = reinterpret_cast<uint16_t *>(cov_smem);                          // [COV_QCAP][COV_TPB]
  uint32_t *bm = reinterpret_cast<uint32_t *>(cov_smem + COV_TPB * COV_QCAP * 2);  // [WIN*WIN/32][COV_TPB]
  const int t = threadIdx.x, b = blockIdx.y, k = blockIdx.x * COV_TPB + t;
  if (blockIdx.x * COV_TPB >= a.count[b]) return;  // whole block beyond the frame's keypoints
  {  // the warp clears all 32 bitmaps together (64 KB, 16-byte stores)
    uint4 *bm4 = reinterpret_cast<uint4 *>(bm);
    for (int i = t; i < COV_TPB * COV_WIN * COV_WIN / 8 / 16; i += COV_TPB) bm4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
  }
  if (k >= a.count[b]) return;
  const size_t px = static_cast<size_t>(a.H) * a.W, ki = static_cast<size_t>(b) * a.cap + k;
  const float *heat = a.heat_inv + b * px;
  int *owner = a.owner + b * px;
  const float *xy = a.kp_xy + ki * 2;
  const int cu = static_cast<int>(xy[0]), cv = static_cast<int>(xy[1]);
  const int ox = cu - COV_WIN / 2, oy = cv - COV_WIN / 2;  // window origin
  const int W = a.W, H = a.H;
  // warm L1 with the keypoint's neighbourhood (a flood is a chain of dependent loads; floods are ~30 pixels)
  for (int r = -8; r <= 8; r++) {
    const int v = min(max(cv + r, 0), H - 1);
    prefetch_l1(heat + v * W + max(cu - 16, 0));
    prefetch_l1(heat + v * W + min(cu + 16, W - 1));
  }
  int head = 0, tail = 0;
  bool over = false;
  sq[(tail++) * COV_TPB + t] = static_cast<uint16_t>((COV_WIN / 2) * COV_WIN + COV_WIN / 2);
  while (head < tail && !over) {
    const int idx = sq[(head++) * COV_TPB + t];
    bm[(idx >> 5) * COV_TPB + t] |= 1u << (idx & 31);  // visited when popped
    const int wx = idx & (COV_WIN - 1), wy = idx / COV_WIN;
    const int u = ox + wx, v = oy + wy, pix = v * W + u;
    atomicMin(owner + pix, k);  // result unused: fire-and-forget reduction
    // neighbours in the reference's order: left, up, right, down; its boundary tests: xx > 0, yy > 0, xx < w, yy < h
    const bool ok[4] = {u - 1 > 0, v - 1 > 0, u + 1 < W, v + 1 < H};
    const int dpix[4] = {-1, -W, 1, W}, didx[4] = {-1, -COV_WIN, 1, COV_WIN};
    const bool inwin[4] = {wx > 0, wy > 0, wx < COV_WIN - 1, wy < COV_WIN - 1};
    const float here = __ldg(heat + pix);
    float hv[4];
#pragma unroll
    for (int d = 0; d < 4; d++) hv[d] = __ldg(heat + (ok[d] ? pix + dpix[d] : pix));  // all loads of the pop in flight together
#pragma unroll
    for (int d = 0; d < 4; d++) {
      if (!ok[d] || !(hv[d] > 0.0f && hv[d] < here)) continue;
      if (!inwin[d]) { over = true; break; }
      const int nidx = idx + didx[d];
      if ((bm[(nidx >> 5) * COV_TPB + t] >> (nidx & 31)) & 1u) continue;
      if (tail >= COV_QCAP) { over = true; break; }
      sq[(tail++) * COV_TPB + t] = static_cast<uint16_t>(nidx);
    }
  }
  if (over) {
    a.frame_flag[b] = 1;
    a.qlen[ki] = -1;
    return;
  }
  uint32_t *q = a.queue + ki * COV_QCAP;
  for (int i = 0; i < tail; i++) {
    const int idx = sq[i * COV_TPB + t];
    q[i] = static_cast<uint32_t>((oy + idx / COV_WIN) * W + ox + (idx & (COV_WIN - 1)));
  }
  a.qlen[ki] = tail;
}

// Phase 2: clean keypoints are final.
__global__ void __launch_bounds__(128) cov_finish_kernel(const CovArgs a) {
  const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.count[b]) return;
  const size_t px = static_cast<size_t>(a.H) * a.W, ki = static_cast<size_t>(b) * a.cap + k;
  const float *heat = a.heat_inv + b * px;
  const int *owner = a.owner + b * px;
  const uint32_t *q = a.queue + ki * COV_QCAP;
  const float *xy = a.kp_xy + ki * 2;
  const int cu = static_cast<int>(xy[0]), cv = static_cast<int>(xy[1]);
  a.response[ki] = heat[cv * a.W + cu];  // kp.response = heat_inv(v, u), sp_extractor.cpp:271
  const int n = a.qlen[ki];
  if (a.frame_flag[b]) { a.qlen[ki] = -1; return; }  // whole frame goes through the sequential path
  int dirty = 0;
  for (int i = 0; i < n; i++) dirty |= (owner[q[i]] != k);
  if (n <= 0 || dirty) {
    a.qlen[ki] = -1;  // replayed by cov_replay_kernel
    return;
  }
  cov_moments(heat, q, n, a.W, cu, cv, a.cov2 + ki * 2, a.cov2_inv + ki * 2);
  uint32_t *visited = a.visited + static_cast<size_t>(b) * a.vis_words;
  for (int i = 0; i < n; i++) atomicOr(visited + (q[i] >> 5), 1u << (q[i] & 31));
}

// Phase 3: sequential replay (one warp per frame) of the keypoints that were not clean, in index order, against the
// true visited map.  A pop is a chain of dependent accesses, so everything it touches is kept close: the queue
// (entries packed v << 16 | u: no divisions) and, for frames up to 524288 pixels, the frame's visited bitmap in
// shared memory; heat_inv on the read-only path.  Lane 0 runs the floods; all lanes gather the dirty list (ordered
// ballot compaction) and compute the per-entry moment terms, lane 0 only performs the ordered additions.
__global__ void __launch_bounds__(32) cov_replay_kernel(const CovArgs a) {
  extern __shared__ uint32_t cov_rsmem[];
  uint32_t *rq = cov_rsmem;                    // [COV_SEQ_QCAP]
  uint32_t *s_vis = cov_rsmem + COV_SEQ_QCAP;  // [COV_SEQ_BITMAP_WORDS]
  __shared__ uint16_t s_dirty[4096];
  __shared__ float s_t0[1024], s_t1[1024];
  const int b = blockIdx.x, lane = threadIdx.x;
  uint32_t *g_vis = a.visited + static_cast<size_t>(b) * a.vis_words;
  const bool in_smem = a.vis_words <= COV_SEQ_BITMAP_WORDS;
  if (in_smem)
    for (int i = lane; i < a.vis_words; i += 32) s_vis[i] = g_vis[i];
  uint32_t *vis = in_smem ? s_vis : g_vis;
  const size_t px = static_cast<size_t>(a.H) * a.W;
  const float *heat = a.heat_inv + b * px;
  const int n_kp = a.count[b], W = a.W, H = a.H;
  // ordered list of the keypoints to replay
  int n_dirty = 0;
  for (int k0 = 0; k0 < n_kp; k0 += 32) {
    const int k = k0 + lane;
    const bool d = k < n_kp && a.qlen[static_cast<size_t>(b) * a.cap + k] < 0;
    const unsigned m = __ballot_sync(0xffffffffu, d);
    if (d && n_dirty + __popc(m & ((1u << lane) - 1)) < 4096) s_dirty[n_dirty + __popc(m & ((1u << lane) - 1))] = static_cast<uint16_t>(k);
    n_dirty += __popc(m);
  }
  __syncwarp();
  if (n_dirty > 4096) {  // cannot happen (cap <= 4096 keypoints are supported by the matcher as well)
    if (lane == 0) atomicExch(a.overflow, 1);
    return;
  }
  int stat_p = 0;
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
      if (lane == 0) atomicExch(a.overflow, 1);
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
      a.qlen[ki] = tail;
    }
    stat_p += tail;
  }
  if (lane == 0) {
    a.n_replay[2 * b] = n_dirty;
    a.n_replay[2 * b + 1] = stat_p;
  }
}

}  // namespace spfe
