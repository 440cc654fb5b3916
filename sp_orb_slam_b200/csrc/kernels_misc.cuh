// CUDA-core kernels around the tensor-core convolutions:
//   conv1a_kernel       u8 -> fp32 3x3 conv (1 -> 64) + ReLU -> fp16 NHWC   (sp_extractor.cpp:81, :386-390)
//   nms_kernel          threshold + exact greedy NMS + cap + border + raster compaction + occ_grid (:122, :161-250, :489-498)
//   sample_desc_kernel  bilinear descriptor sampling + L2 normalise at the survivors (:134-148)
//   heat_norm_kernel    to_heat min/max normalisation (:461-474)
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "conv_tc.cuh"

namespace spfe {

// ---------------------------------------------------------------------------
// conv1a: K = 9 is not a tensor-core shape and this layer is the precision-
// critical one (|w| up to 197, SURVEY.md §7 hard part 1): fp32 FFMA on CUDA
// cores, input scaled exactly like cv::Mat::convertTo(CV_32F, 1/255).
// Thread = (pixel slot, 8-channel group): the 9x8 weights of the group live in
// registers for the whole 64x16-pixel tile, inputs come from a shared patch
// (broadcast reads), and a warp stores 4 pixels x 128 B = 512 contiguous bytes
// per instruction.  72 FFMA per 9 LDS + 1 STG.128.
// ---------------------------------------------------------------------------
// XP (exact mode): the fp32 result is stored as a hi + lo fp16 pair, 128 channels per pixel ([hi 64 | lo 64]).
constexpr int C1A_TW = 64, C1A_TH = 16;
template <bool XP>
__global__ void __launch_bounds__(256) conv1a_kernel(const uint8_t *__restrict__ in, __half *__restrict__ out,
                                                     const float *__restrict__ wgt /*[9][64]*/,
                                                     const float *__restrict__ bias /*[64]*/, int B, int H, int W) {
  __shared__ float s_in[C1A_TH + 2][C1A_TW + 2];
  const int g = threadIdx.x & 7, slot = threadIdx.x >> 3;  // channel group, pixel slot 0..31
  const int x0 = blockIdx.x * C1A_TW, y0 = blockIdx.y * C1A_TH, b = blockIdx.z;
  const uint8_t *img = in + static_cast<size_t>(b) * H * W;
  float2 w[9][4], bs[4];  // channel pairs: FFMA2 does two fp32 FMAs per instruction
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const float4 lo = *reinterpret_cast<const float4 *>(wgt + t * 64 + g * 8);
    const float4 hi = *reinterpret_cast<const float4 *>(wgt + t * 64 + g * 8 + 4);
    w[t][0] = make_float2(lo.x, lo.y); w[t][1] = make_float2(lo.z, lo.w);
    w[t][2] = make_float2(hi.x, hi.y); w[t][3] = make_float2(hi.z, hi.w);
  }
#pragma unroll
  for (int c = 0; c < 4; c++) bs[c] = make_float2(bias[g * 8 + 2 * c], bias[g * 8 + 2 * c + 1]);
  const float scale = 1.0f / 255.0f;
  for (int i = threadIdx.x; i < (C1A_TH + 2) * (C1A_TW + 2); i += 256) {
    const int r = i / (C1A_TW + 2), c = i - r * (C1A_TW + 2);
    const int y = y0 + r - 1, x = x0 + c - 1;
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) v = static_cast<float>(img[static_cast<size_t>(y) * W + x]) * scale;
    s_in[r][c] = v;
  }
  __syncthreads();
#pragma unroll 2
  for (int p = slot; p < C1A_TW * C1A_TH; p += 32) {
    const int ty = p / C1A_TW, tx = p - ty * C1A_TW;
    const int x = x0 + tx, y = y0 + ty;
    float2 acc[4];
#pragma unroll
    for (int c = 0; c < 4; c++) acc[c] = bs[c];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const float xin = s_in[ty + t / 3][tx + t % 3];
      const float2 x2 = make_float2(xin, xin);
#pragma unroll
      for (int c = 0; c < 4; c++) ffma2(acc[c], w[t][c], x2);
    }
    if (x < W && y < H) {
      if constexpr (XP) {
        uint4 o, l;
        split_h2(fmaxf(acc[0].x, 0.f), fmaxf(acc[0].y, 0.f), o.x, l.x);
        split_h2(fmaxf(acc[1].x, 0.f), fmaxf(acc[1].y, 0.f), o.y, l.y);
        split_h2(fmaxf(acc[2].x, 0.f), fmaxf(acc[2].y, 0.f), o.z, l.z);
        split_h2(fmaxf(acc[3].x, 0.f), fmaxf(acc[3].y, 0.f), o.w, l.w);
        __half *d = out + ((static_cast<size_t>(b) * H + y) * W + x) * 128 + g * 8;
        *reinterpret_cast<uint4 *>(d) = o;
        *reinterpret_cast<uint4 *>(d + 64) = l;
      } else {
      uint4 o;
      o.x = pack_h2(fmaxf(acc[0].x, 0.f), fmaxf(acc[0].y, 0.f));
      o.y = pack_h2(fmaxf(acc[1].x, 0.f), fmaxf(acc[1].y, 0.f));
      o.z = pack_h2(fmaxf(acc[2].x, 0.f), fmaxf(acc[2].y, 0.f));
      o.w = pack_h2(fmaxf(acc[3].x, 0.f), fmaxf(acc[3].y, 0.f));
      *reinterpret_cast<uint4 *>(out + ((static_cast<size_t>(b) * H + y) * W + x) * 64 + g * 8) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// NMS.  The reference sorts candidates by score and greedily suppresses a
// (2r+1)^2 window around each survivor.  There is at most one candidate per
// 8x8 cell and r <= 8, so a candidate only interacts with the 8 neighbouring
// cells and the greedy result is the unique fixpoint of
//   kept(c)  <=>  every higher-priority neighbour within Chebyshev r is suppressed
// (priority = score desc, ties by raster cell index = stable sort order).
// The cap "stop once more than nf are kept" keeps the nf+1 best survivors; it
// is applied BEFORE the border filter, and the output is in raster pixel order
// with occ_grid[cell] = output index.  One CTA per frame.
// ---------------------------------------------------------------------------
struct NmsArgs {
  const float *score;     // [B][cells]
  const uint8_t *argmax;  // [B][cells]
  int hc, wc;             // cells per column / row
  float thresh;
  int radius, border, cap;  // cap = max_keypoints + 1
  int *count;               // [B]
  float *kp_xy;             // [B][cap][2]
  float *kp_score;          // [B][cap]
  int16_t *occ;             // [B][cells]
  unsigned long long *scratch;  // [B][cells] key list
  int list_smem;                // 1: the key list lives behind the cell arrays in dynamic shared memory (cells * 8 more bytes)
};

enum { ST_NONE = 0, ST_UNDEC = 1, ST_KEPT = 2, ST_SUPP = 3 };

__global__ void __launch_bounds__(1024) nms_kernel(const NmsArgs p) {
  extern __shared__ uint8_t nms_smem[];
  const int cells = p.hc * p.wc;
  float *s_sc = reinterpret_cast<float *>(nms_smem);
  uint8_t *s_pos = reinterpret_cast<uint8_t *>(s_sc + cells);
  uint8_t *s_st = s_pos + cells;
  uint8_t *s_new = s_st + cells;  // next state of the undecided cells (Jacobi iteration: reads and writes never race)
  __shared__ int s_cnt, s_cnt2;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float *score = p.score + static_cast<size_t>(b) * cells;
  const uint8_t *amax = p.argmax + static_cast<size_t>(b) * cells;
  int16_t *occ = p.occ + static_cast<size_t>(b) * cells;
  // key lists of the two counting sorts: in shared memory when the launch provided room for them (p.list_smem), else in global scratch
  unsigned long long *list = p.list_smem ? reinterpret_cast<unsigned long long *>(nms_smem + ((cells * 7 + 7) & ~7))
                                         : p.scratch + static_cast<size_t>(b) * cells;
  const int W = p.wc * 8, H = p.hc * 8;

  if (tid == 0) { s_cnt = 0; s_cnt2 = 0; }
  __syncthreads();
  // The undecided candidates live in a compact list (two halves of the frame's scratch, ping-pong): a round touches only
  // the cells that are still open -- a quarter of the cells in round 0 on dense frames, a handful after three rounds --
  // instead of striding over every cell (1920x1080: 32 400 cells per round and frame on one SM).
  int *ul_cur = reinterpret_cast<int *>(p.scratch + static_cast<size_t>(b) * cells), *ul_nxt = ul_cur + cells;
  const unsigned lane_lt0 = (1u << (tid & 31)) - 1u;
  for (int c0 = 0; c0 < cells; c0 += nt) {
    const int c = c0 + tid;
    bool cand = false;
    if (c < cells) {
      const float sc = score[c];
      cand = sc >= p.thresh;
      s_sc[c] = sc;
      s_pos[c] = amax[c];
      s_st[c] = cand ? ST_UNDEC : ST_NONE;
      occ[c] = -1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, cand);
    int base = 0;
    if ((tid & 31) == 0 && m) base = atomicAdd(&s_cnt, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (cand) ul_cur[base + __popc(m & lane_lt0)] = c;
  }
  __syncthreads();

  // ---- fixpoint of the greedy suppression (Jacobi: a round reads only the previous round's states)
  int U = s_cnt;
  for (int round = 0; round < cells + 2 && U > 0; round++) {
    for (int e = tid; e < U; e += nt) {
      const int c = ul_cur[e];
      const int cy = c / p.wc, cx = c - cy * p.wc;
      const int px = cx * 8 + (s_pos[c] & 7), py = cy * 8 + (s_pos[c] >> 3);
      const float sc = s_sc[c];
      bool any_kept = false, all_supp = true;
      for (int dy = -1; dy <= 1; dy++) {
        const int ny = cy + dy;
        if (ny < 0 || ny >= p.hc) continue;
        for (int dx = -1; dx <= 1; dx++) {
          const int nx = cx + dx;
          if ((dx == 0 && dy == 0) || nx < 0 || nx >= p.wc) continue;
          const int d = ny * p.wc + nx;
          const int st = s_st[d];
          if (st == ST_NONE || st == ST_SUPP) continue;
          const int qx = nx * 8 + (s_pos[d] & 7), qy = ny * 8 + (s_pos[d] >> 3);
          if (abs(qx - px) > p.radius || abs(qy - py) > p.radius) continue;
          const float sd = s_sc[d];
          if (sd > sc || (sd == sc && d < c)) {
            if (st == ST_KEPT) any_kept = true;
            else all_supp = false;
          }
        }
      }
      s_new[c] = any_kept ? ST_SUPP : (all_supp ? ST_KEPT : ST_UNDEC);
    }
    if (tid == 0) s_cnt2 = 0;
    __syncthreads();
    for (int e0 = 0; e0 < U; e0 += nt) {  // apply, and keep the still-undecided cells for the next round
      const int e = e0 + tid;
      bool open = false;
      int c = 0;
      if (e < U) {
        c = ul_cur[e];
        const uint8_t st = s_new[c];
        s_st[c] = st;
        open = st == ST_UNDEC;
      }
      const unsigned m = __ballot_sync(0xffffffffu, open);
      int base = 0;
      if ((tid & 31) == 0 && m) base = atomicAdd(&s_cnt2, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (open) ul_nxt[base + __popc(m & lane_lt0)] = c;
    }
    __syncthreads();
    U = s_cnt2;
    int *t = ul_cur; ul_cur = ul_nxt; ul_nxt = t;
    __syncthreads();  // everybody has read s_cnt2 before the next round clears it
  }
  if (tid == 0) { s_cnt = 0; s_cnt2 = 0; }
  __syncthreads();

  // ---- cap: keep the `cap` best survivors (score desc, ties by cell index asc)
  // (list appends are warp-aggregated: one shared-memory atomic per warp instead of one per survivor)
  const unsigned lane_lt = (1u << (tid & 31)) - 1u;
  for (int c0 = 0; c0 < cells; c0 += nt) {
    const int c = c0 + tid;
    const bool kept = c < cells && s_st[c] == ST_KEPT;
    const unsigned m = __ballot_sync(0xffffffffu, kept);
    int base = 0;
    if ((tid & 31) == 0 && m) base = atomicAdd(&s_cnt, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (kept)
      list[base + __popc(m & lane_lt)] = (static_cast<unsigned long long>(__float_as_uint(s_sc[c])) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(c));
  }
  __syncthreads();
  const int K = s_cnt;
  if (K > p.cap) {
    // the cap-th largest key by radix selection (8 passes over the K keys, one byte each): O(K) instead of the O(K^2)
    // rank-by-counting that cost 0.1 ms per 1920x1080 frame (K ~ 5300 survivors against a budget of 2001)
    __shared__ int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_want;
    if (tid == 0) { s_prefix = 0ull; s_want = p.cap; }
    for (int shift = 56; shift >= 0; shift -= 8) {
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned long long hi_mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
      for (int e = tid; e < K; e += nt) {
        const unsigned long long key = list[e];
        if ((key & hi_mask) == prefix) atomicAdd(&s_hist[static_cast<int>((key >> shift) & 255ull)], 1);
      }
      __syncthreads();
      if (tid < 32) {  // digit d of the threshold: the largest d with  #(keys with a digit >= d)  >= want
        int cnt[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { cnt[j] = s_hist[255 - (tid * 8 + j)]; sum += cnt[j]; }  // lane 0 owns digits 255..248
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += v;
        }
        const int want = s_want;
        int above = incl - sum;  // keys with a larger digit than this lane's eight
        const bool mine = above < want && incl >= want;
        if (mine) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (above + cnt[j] >= want) {
              s_prefix = prefix | (static_cast<unsigned long long>(255 - (tid * 8 + j)) << shift);
              s_want = want - above;
              break;
            }
            above += cnt[j];
          }
        }
      }
      __syncthreads();
    }
    const unsigned long long cut = s_prefix;  // == the cap-th largest key (keys are unique: the cell index is part of them)
    for (int e = tid; e < K; e += nt) {
      const unsigned long long key = list[e];
      if (key < cut) s_st[0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFu)] = ST_SUPP;
    }
  }
  __syncthreads();

  // ---- border filter + raster (v outer, u inner) ordering
  for (int c0 = 0; c0 < cells; c0 += nt) {
    const int c = c0 + tid;
    bool in = false;
    int px = 0, py = 0;
    if (c < cells && s_st[c] == ST_KEPT) {
      const int cy = c / p.wc, cx = c - cy * p.wc;
      px = cx * 8 + (s_pos[c] & 7);
      py = cy * 8 + (s_pos[c] >> 3);
      in = px >= p.border && px < W - p.border && py >= p.border && py < H - p.border;
    }
    const unsigned m = __ballot_sync(0xffffffffu, in);
    int base = 0;
    if ((tid & 31) == 0 && m) base = atomicAdd(&s_cnt2, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (in) list[base + __popc(m & lane_lt)] = (static_cast<unsigned long long>(py * W + px) << 32) | static_cast<unsigned>(c);
  }
  __syncthreads();
  const int Nk = s_cnt2;
  float *kp_xy = p.kp_xy + static_cast<size_t>(b) * p.cap * 2;
  float *kp_sc = p.kp_score + static_cast<size_t>(b) * p.cap;
  for (int e = tid; e < Nk; e += nt) {
    const unsigned long long key = list[e];
    int rank = 0;
    for (int f = 0; f < Nk; f++) rank += (list[f] < key);
    const int c = static_cast<int>(key & 0xFFFFFFFFu);
    const int pix = static_cast<int>(key >> 32);
    kp_xy[2 * rank] = static_cast<float>(pix % W);
    kp_xy[2 * rank + 1] = static_cast<float>(pix / W);
    kp_sc[rank] = s_sc[c];
    occ[c] = static_cast<int16_t>(rank);
  }
  if (tid == 0) p.count[b] = Nk;
}

// ---------------------------------------------------------------------------
// Descriptor sampling at the NMS survivors: grid_sampler_2d(coarse, pts,
// bilinear, zeros, align_corners=true) then per-keypoint L2 normalisation.
// The reference samples every candidate and discards most of them in nms();
// sampling only survivors gives the same rows.  One warp per keypoint, each
// lane owns 8 of the 256 channels (one 16-byte load per corner).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_desc_kernel(const __half *__restrict__ coarse /*[B][hc][wc][256]*/,
                                                          const float *__restrict__ kp_xy, const int *__restrict__ count,
                                                          float *__restrict__ desc /*[B][cap][256]*/, int hc, int wc,
                                                          int cap, __half *__restrict__ x16 /*[B][rows_pad][256] or null*/,
                                                          int rows_pad) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= count[b]) return;
  const float x = kp_xy[(static_cast<size_t>(b) * cap + i) * 2], y = kp_xy[(static_cast<size_t>(b) * cap + i) * 2 + 1];
  const float W = static_cast<float>(wc * 8), H = static_cast<float>(hc * 8);
  // sp_extractor.cpp:137-138 then ATen grid_sampler_unnormalize(align_corners=true)
  const float xs = __fsub_rn(__fdiv_rn(x, W * 0.5f), 1.0f), ys = __fsub_rn(__fdiv_rn(y, H * 0.5f), 1.0f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(xs, 1.0f), 2.0f), static_cast<float>(wc - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(ys, 1.0f), 2.0f), static_cast<float>(hc - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix, wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
  const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __half *base = coarse + static_cast<size_t>(b) * hc * wc * 256 + lane * 8;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
    if (xx < 0 || xx >= wc || yy < 0 || yy >= hc) continue;
    const uint4 raw = *reinterpret_cast<const uint4 *>(base + (static_cast<size_t>(yy) * wc + xx) * 256);
    const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float2 f = __half22float2(h2[j]);
      acc[2 * j] = fmaf(wgt[k], f.x, acc[2 * j]);
      acc[2 * j + 1] = fmaf(wgt[k], f.y, acc[2 * j + 1]);
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; j++) ss = fmaf(acc[j], acc[j], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = sqrtf(ss);
  float4 *dst = reinterpret_cast<float4 *>(desc + (static_cast<size_t>(b) * cap + i) * 256 + lane * 8);
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = acc[j] / nrm;
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  if (x16 != nullptr) {  // fp16 copy feeding the tensor-core matcher (candidate generation only)
    uint4 o;
    o.x = pack_h2(acc[0], acc[1]);
    o.y = pack_h2(acc[2], acc[3]);
    o.z = pack_h2(acc[4], acc[5]);
    o.w = pack_h2(acc[6], acc[7]);
    *reinterpret_cast<uint4 *>(x16 + (static_cast<size_t>(b) * rows_pad + i) * 256 + lane * 8) = o;
  }
}

// ---------------------------------------------------------------------------
// to_heat: heat = (-x - min)/(max - min), heat_inv = (max + x)/(max - min)
// with min/max of -x, folded the way OpenCV's MatExpr evaluates it (one
// scale+shift per output, coefficients formed in double then narrowed).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) heat_norm_kernel(const float *__restrict__ heat_log, const unsigned *__restrict__ minmax,
                                                        float *__restrict__ heat, float *__restrict__ heat_inv,
                                                        float *__restrict__ minmax_out, int n_per_frame) {
  __shared__ float s_c[4];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    const float lo = f32_from_ordered(minmax[2 * b]), hi = f32_from_ordered(minmax[2 * b + 1]);
    const double mn = static_cast<double>(-hi), mx = static_cast<double>(-lo);  // of img = -heat_log
    const double inv = 1.0 / (mx - mn);
    s_c[0] = static_cast<float>(-1.0 * inv);
    s_c[1] = static_cast<float>((-mn) * inv);
    s_c[2] = static_cast<float>(1.0 * inv);
    s_c[3] = static_cast<float>(mx * inv);
    if (blockIdx.x == 0 && minmax_out != nullptr) {
      minmax_out[2 * b] = static_cast<float>(mn);
      minmax_out[2 * b + 1] = static_cast<float>(mx);
    }
  }
  __syncthreads();
  const float a0 = s_c[0], b0 = s_c[1], a1 = s_c[2], b1 = s_c[3];
  const size_t off = static_cast<size_t>(b) * n_per_frame;
  const float4 *src = reinterpret_cast<const float4 *>(heat_log + off);
  float4 *d0 = reinterpret_cast<float4 *>(heat + off), *d1 = reinterpret_cast<float4 *>(heat_inv + off);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_per_frame / 4; i += gridDim.x * blockDim.x) {
    const float4 v = src[i];
    if (heat != nullptr)  // heat_ is only materialised when it is copied to the host (SPFE_EMIT_HEAT)
      d0[i] = make_float4(__fadd_rn(__fmul_rn(v.x, a0), b0), __fadd_rn(__fmul_rn(v.y, a0), b0),
                          __fadd_rn(__fmul_rn(v.z, a0), b0), __fadd_rn(__fmul_rn(v.w, a0), b0));
    d1[i] = make_float4(__fadd_rn(__fmul_rn(v.x, a1), b1), __fadd_rn(__fmul_rn(v.y, a1), b1),
                        __fadd_rn(__fmul_rn(v.z, a1), b1), __fadd_rn(__fmul_rn(v.w, a1), b1));
  }
}

}  // namespace spfe
