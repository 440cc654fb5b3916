// Brute-force mutual nearest-neighbour matcher over 256-d float descriptors.
//
// Replaces cv::BFMatcher(NORM_L2, crossCheck=true)::match as called by
// SPMatcher::SearchByBruteForce (reference orb_slam2/src/cv/sp_matcher.cpp:1666-1669,
// orb_slam2/src/cv/sp_matcher_loop.cpp:365-368):
//   q2t[i] = j*  where j* = first-index argmin_j ||q_i - t_j||  and  first-index argmin_i ||q_i - t_j*|| == i
// Squared distances are accumulated in fp32 exactly as sum_k (q_ik - t_jk)^2;
// each 64x64 tile of the distance matrix is reduced to per-row / per-column
// minima which are merged across tiles with 64-bit atomicMin on
// (float_bits(d2) << 32 | index) -- smaller distance wins, ties go to the lower
// index, which is BFMatcher's first-index rule.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace spfe {

struct MatchScratch {
  unsigned long long *rowbest = nullptr, *colbest = nullptr;  // [cap]
  int *q2t = nullptr;                                          // [cap]
  float *dist = nullptr;                                       // [cap]
  int *dn = nullptr;                                           // [2] nq, nt
  int cap = 0;
};

__global__ void init_minmax_kernel(unsigned *mm, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    mm[2 * b] = 0xFFFFFFFFu;
    mm[2 * b + 1] = 0u;
  }
}

__global__ void match_init_kernel(unsigned long long *rowbest, unsigned long long *colbest, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    rowbest[i] = ~0ull;
    colbest[i] = ~0ull;
  }
}

// One matching problem per blockIdx.z.  Problem z matches query set z against
// train set z-1; train set of problem 0 is the "carry" (last frame of the
// previous batch of the same camera stream), or an explicit set for the
// host-pointer API (q_stride == 0 && z == 0).
struct MatchArgs {
  const float *q;        // [Z][cap][256] query descriptors of frame z
  const int *nq;         // [Z]
  const float *t0;       // train set of problem 0
  const int *nt0;        // its row count
  unsigned long long *rowbest, *colbest;  // [Z][cap]
  int *q2t;              // [Z][cap]
  float *dist;           // [Z][cap]
  int cap;
  // k-NN pass 2 (spfe_match_knn2): skip each row's best column (low word of excl[i]), write the row minima of the rest
  // into rowbest, leave the column minima alone
  const unsigned long long *excl;
};
__device__ __forceinline__ void match_problem(const MatchArgs &a, int z, const float *&q, const float *&t, int &nq, int &nt) {
  q = a.q + static_cast<size_t>(z) * a.cap * 256;
  nq = a.nq[z];
  if (z == 0) { t = a.t0; nt = *a.nt0; }
  else { t = a.q + static_cast<size_t>(z - 1) * a.cap * 256; nt = a.nq[z - 1]; }
}

// grid (ceil(cap/64) train tiles, ceil(cap/64) query tiles, Z), 256 threads, 4x4 outputs per thread.
__global__ void __launch_bounds__(256) match_dist_kernel(const MatchArgs ma) {
  const float *q, *t;
  int nq, nt;
  match_problem(ma, blockIdx.z, q, t, nq, nt);
  unsigned long long *rowbest = ma.rowbest + static_cast<size_t>(blockIdx.z) * ma.cap;
  unsigned long long *colbest = ma.colbest + static_cast<size_t>(blockIdx.z) * ma.cap;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  if (i0 >= nq || j0 >= nt) return;
  __shared__ float sq[16][68], st[16][68];
  __shared__ unsigned long long s_col[16][64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;  // loader: row 0..63, k offset 0,4,8,12
  for (int k0 = 0; k0 < 256; k0 += 16) {
    float4 vq = make_float4(0, 0, 0, 0), vt = make_float4(0, 0, 0, 0);
    if (i0 + lr < nq) vq = *reinterpret_cast<const float4 *>(q + static_cast<size_t>(i0 + lr) * 256 + k0 + lk);
    if (j0 + lr < nt) vt = *reinterpret_cast<const float4 *>(t + static_cast<size_t>(j0 + lr) * 256 + k0 + lk);
    __syncthreads();
    sq[lk + 0][lr] = vq.x; sq[lk + 1][lr] = vq.y; sq[lk + 2][lr] = vq.z; sq[lk + 3][lr] = vq.w;
    st[lk + 0][lr] = vt.x; st[lk + 1][lr] = vt.y; st[lk + 2][lr] = vt.z; st[lk + 3][lr] = vt.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; r++) { a[r] = sq[k][ty * 4 + r]; b[r] = st[k][tx * 4 + r]; }
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int cidx = 0; cidx < 4; cidx++) {
          const float d = a[r] - b[cidx];
          acc[r][cidx] = fmaf(d, d, acc[r][cidx]);
        }
    }
  }
  // row minima: reduce this thread's 4 columns, then across the 16 tx lanes (same half-warp)
#pragma unroll
  for (int r = 0; r < 4; r++) {
    unsigned long long best = ~0ull;
    unsigned skip = 0xFFFFFFFFu;
    if (ma.excl != nullptr && i0 + ty * 4 + r < nq)
      skip = static_cast<unsigned>(ma.excl[static_cast<size_t>(blockIdx.z) * ma.cap + i0 + ty * 4 + r] & 0xFFFFFFFFu);
#pragma unroll
    for (int cidx = 0; cidx < 4; cidx++) {
      const int j = j0 + tx * 4 + cidx;
      if (j < nt && static_cast<unsigned>(j) != skip) {
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(acc[r][cidx])) << 32) | static_cast<unsigned>(j);
        best = key < best ? key : best;
      }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    const int i = i0 + ty * 4 + r;
    if (tx == 0 && i < nq) atomicMin(rowbest + i, best);
  }
  if (ma.excl != nullptr) return;  // block-uniform
  // column minima: per-thread over its 4 rows, then across the 16 ty groups through shared memory
#pragma unroll
  for (int cidx = 0; cidx < 4; cidx++) {
    unsigned long long best = ~0ull;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int i = i0 + ty * 4 + r;
      if (i < nq) {
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(acc[r][cidx])) << 32) | static_cast<unsigned>(i);
        best = key < best ? key : best;
      }
    }
    s_col[ty][tx * 4 + cidx] = best;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    unsigned long long best = ~0ull;
#pragma unroll
    for (int g = 0; g < 16; g++) {
      const unsigned long long v = s_col[g][threadIdx.x];
      best = v < best ? v : best;
    }
    const int j = j0 + threadIdx.x;
    if (j < nt) atomicMin(colbest + j, best);
  }
}

__global__ void match_final_kernel(const MatchArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, cap = a.cap;
  if (i >= cap) return;
  const size_t off = static_cast<size_t>(blockIdx.y) * cap;
  const unsigned long long *rowbest = a.rowbest + off, *colbest = a.colbest + off;
  int *q2t = a.q2t + off;
  float *dist = a.dist + off;
  int out = -1;
  float d = 0.f;
  if (i < a.nq[blockIdx.y]) {
    const unsigned long long rb = rowbest[i];
    if (rb != ~0ull) {
      const int j = static_cast<int>(rb & 0xFFFFFFFFu);
      d = sqrtf(__uint_as_float(static_cast<unsigned>(rb >> 32)));
      if (static_cast<int>(colbest[j] & 0xFFFFFFFFu) == i) out = j;
    }
  }
  q2t[i] = out;
  dist[i] = d;
}


// Exact 2-NN of every query row: best / second-best keys -> idx[i][2], dist[i][2] (-1 / 0 where there is no such row).
__global__ void knn2_final_kernel(const unsigned long long *first, const unsigned long long *second, int nq, int *idx, float *dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long k[2] = {first[i], second[i]};
#pragma unroll
  for (int r = 0; r < 2; r++) {
    idx[2 * i + r] = k[r] == ~0ull ? -1 : static_cast<int>(k[r] & 0xFFFFFFFFu);
    dist[2 * i + r] = k[r] == ~0ull ? 0.f : sqrtf(__uint_as_float(static_cast<unsigned>(k[r] >> 32)));
  }
}

// ---------------------------------------------------------------------------
// Tensor-core path for the in-pipeline stream matching (SPFE_MATCH_PREV).
// conv_tc_kernel<EPI_TOP2> computes fp16 dot products Q.T^T on the tensor core
// and keeps, per row and per 256-column block, the two best candidates.  The
// descriptors are unit vectors, so ranking by dot product == ranking by L2; an
// fp16 dot product is within 2^-10 of the exact one, so every candidate whose
// score is within MATCH_MARGIN of the row's best is re-ranked here with the
// exact fp32 squared distance (ties -> lower index, BFMatcher's rule).
// One warp per row; grid (ceil(cap / 8), Z pairs, 2 directions).
// ---------------------------------------------------------------------------
// Why the result is provably the exact one.  Descriptors are unit vectors; rounding both operands to fp16 changes a
// dot product by at most 2 * 2^-11 * sum |q_k t_k| <= 2^-10 (Cauchy-Schwarz), the fp32 tensor-core accumulation and the
// 8 dropped key bits add < 5e-5: |observed - exact| <= beta = 1.03e-3.  Let s_R be the R-th largest observed score of a
// row over all columns (R = 1: nearest neighbour, R = 2: two nearest).  A column of the true top R has an observed
// score >= s_R - 2 beta, and the top R observed columns are among the K >= R nominees of their 256-column blocks, so
// s_R is known from the nominees.  A true top-R column can be missing from the nominees only if its block holds K
// other columns at or above it, i.e. the block's K-th nominee is >= s_R - 2 beta: such a block is re-scanned
// completely in fp32; everywhere else the nominees within MATCH_MARGIN (>= 2 beta) of s_R contain the answer.  "Block"
// is a VIRTUAL block: the 64 columns of a 256-column GEMM tile with the same residue mod 4 (the epilogue's four
// independent chains, kept apart), so that K near-ties must share a residue class to force a 64-column re-scan.  All survivors are ranked by the exact fp32 squared distance, ties to the
// lower index (BFMatcher's first-index rule).
constexpr float MATCH_MARGIN = 4e-3f;
constexpr float MATCH_RESCAN = 2.1e-3f;  // >= 2 beta

// exact fp32 squared distance of two 256-d rows held 8 dimensions per lane (all lanes return the sum)
__device__ __forceinline__ float warp_dist2(const float4 &m0, const float4 &m1, const float *other, int lane) {
  const float4 o0 = *reinterpret_cast<const float4 *>(other + lane * 8), o1 = *reinterpret_cast<const float4 *>(other + lane * 8 + 4);
  float d, acc = 0.f;
  d = m0.x - o0.x; acc = fmaf(d, d, acc);
  d = m0.y - o0.y; acc = fmaf(d, d, acc);
  d = m0.z - o0.z; acc = fmaf(d, d, acc);
  d = m0.w - o0.w; acc = fmaf(d, d, acc);
  d = m1.x - o1.x; acc = fmaf(d, d, acc);
  d = m1.y - o1.y; acc = fmaf(d, d, acc);
  d = m1.z - o1.z; acc = fmaf(d, d, acc);
  d = m1.w - o1.w; acc = fmaf(d, d, acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// One warp re-ranks one row: `c` = its [NV][K] nominees (observed score, column) of the NV = 4 * NB virtual blocks (the
// 64 columns of a 256-column block with the same residue mod 4), `other` = the fp32 rows of the other set (n_other of
// them).  best[0 .. R) receive the R smallest keys (bits(d^2) << 32 | column), ~0 where none exists.
template <int K, int R>
__device__ __forceinline__ void rerank_row(const float *me, const float *other, int n_other, const float2 *c, int NB, int lane,
                                           unsigned long long (&best)[R]) {
  static_assert(K >= R && R >= 1 && R <= 2, "nominees per block must cover the ranks asked for");
  const int total = NB * 4 * K;
  // best / second-best observed score among the nominees: lanes scan disjoint entries, then a butterfly merge
  float s1 = -INFINITY, s2 = -INFINITY;
  for (int e = lane; e < total; e += 32) {
    const float x = c[e].x;
    if (x > s1) { s2 = s1; s1 = x; }
    else if (x > s2) s2 = x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t1 = __shfl_xor_sync(0xffffffffu, s1, o), t2 = __shfl_xor_sync(0xffffffffu, s2, o);
    s2 = fmaxf(fminf(s1, t1), fmaxf(s2, t2));
    s1 = fmaxf(s1, t1);
  }
#pragma unroll
  for (int r = 0; r < R; r++) best[r] = ~0ull;
  if (s1 == -INFINITY) return;
  const float s_r = R == 1 ? s1 : (s2 == -INFINITY ? s1 : s2);
  const float thr = s_r - MATCH_MARGIN;        // nominees re-ranked exactly (generous: costs one row read each)
  const float thr_scan = s_r - MATCH_RESCAN;   // blocks re-scanned (tight: 2 beta is what the proof needs)
  const float4 m0 = *reinterpret_cast<const float4 *>(me + lane * 8), m1 = *reinterpret_cast<const float4 *>(me + lane * 8 + 4);
  auto insert = [&](float d2, int idx) {
    unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) | static_cast<unsigned>(idx);
#pragma unroll
    for (int r = 0; r < R; r++)
      if (key < best[r]) { const unsigned long long t = best[r]; best[r] = key; key = t; }
  };
  constexpr int CH = (32 / K) * K;  // entries per step: whole virtual blocks only
  for (int base = 0; base < total; base += CH) {
    const int e = base + lane;
    const bool live = lane < CH && e < total;
    const float2 ce = live ? c[e] : make_float2(-INFINITY, __int_as_float(-1));
    const int idx = __float_as_int(ce.y);
    // K near-ties in one virtual block (its K-th nominee is within 2 beta of s_R): a (K+1)-th may hide behind them
    const unsigned scan = __ballot_sync(0xffffffffu, live && (lane % K) == K - 1 && idx >= 0 && ce.x >= thr_scan);
    const bool block_scanned = lane < CH && ((scan >> ((lane / K) * K + K - 1)) & 1u);  // (lanes >= CH are not live)
    unsigned nom = __ballot_sync(0xffffffffu, live && idx >= 0 && ce.x >= thr && !block_scanned);
    while (nom) {  // warp-uniform
      const int l = __ffs(nom) - 1;
      nom &= nom - 1;
      const int j = __shfl_sync(0xffffffffu, idx, l);
      insert(warp_dist2(m0, m1, other + static_cast<size_t>(j) * 256, lane), j);
    }
    unsigned sc = scan;
    while (sc) {
      const int l = __ffs(sc) - 1;
      sc &= sc - 1;
      const int v = (base + l) / K;  // virtual block: 256-column block v >> 2, residue v & 3
      const int j1 = min(n_other, (v >> 2) * 256 + 256);
      // four columns per step: their row loads and shuffle reductions overlap (the loop is latency-bound otherwise)
      for (int j = (v >> 2) * 256 + (v & 3); j < j1; j += 16) {
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int ju = min(j + 4 * u, j1 - 1);  // clamped duplicates are skipped below
          const float *o = other + static_cast<size_t>(ju) * 256 + lane * 8;
          const float4 o0 = *reinterpret_cast<const float4 *>(o), o1 = *reinterpret_cast<const float4 *>(o + 4);
          float d, a = 0.f;
          d = m0.x - o0.x; a = fmaf(d, d, a);
          d = m0.y - o0.y; a = fmaf(d, d, a);
          d = m0.z - o0.z; a = fmaf(d, d, a);
          d = m0.w - o0.w; a = fmaf(d, d, a);
          d = m1.x - o1.x; a = fmaf(d, d, a);
          d = m1.y - o1.y; a = fmaf(d, d, a);
          d = m1.z - o1.z; a = fmaf(d, d, a);
          d = m1.w - o1.w; a = fmaf(d, d, a);
          acc[u] = a;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < 4; u++) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (j + 4 * u < j1) insert(acc[u], j + 4 * u);
      }
    }
  }
}

struct RerankArgs {
  const float *desc_all;   // [slots][cap][256] fp32 descriptors, slot 0 = carry, slot z+1 = frame z
  const int *count_all;    // [slots]
  const float2 *cand;      // [2][Z][rows_pad][NB][2]
  unsigned long long *rowbest, *colbest;  // [Z][cap]
  int cap, rows_pad, NB, Z;
};

// in-pipeline stream matching (SPFE_MATCH_PREV): one warp per row; grid (ceil(cap / 8), Z pairs, 2 directions)
__global__ void __launch_bounds__(256) match_rerank_kernel(const RerankArgs a) {
  const int z = blockIdx.y, dir = blockIdx.z;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int self_slot = z + 1 - dir, other_slot = z + dir;
  if (row >= a.count_all[self_slot]) return;
  const float2 *c = a.cand + ((static_cast<size_t>(dir) * a.Z + z) * a.rows_pad + row) * a.NB * 4 * 2;
  unsigned long long best[1];
  rerank_row<2, 1>(a.desc_all + (static_cast<size_t>(self_slot) * a.cap + row) * 256, a.desc_all + static_cast<size_t>(other_slot) * a.cap * 256,
                   a.count_all[other_slot], c, a.NB, lane, best);
  if (lane == 0) (dir == 0 ? a.rowbest : a.colbest)[static_cast<size_t>(z) * a.cap + row] = best[0];
}

// descriptor-set matching (spfe_match_mutual_nn / spfe_match_knn2 and their *_sets forms): one direction per launch,
// three nominees per block, exact best (R = 1) or best two (R = 2) of every row of set A among the rows of set B
struct SetRerankArgs {
  const float *a_rows, *b_rows;  // fp32 [n][256]
  int n_a, n_b, rows_pad, NB, dir;
  const float2 *cand;            // [2][1][rows_pad][NB][3]
  unsigned long long *out1, *out2;  // [n_a] best key (and second best for R = 2)
};
template <int R>
__global__ void __launch_bounds__(256) match_rerank_set_kernel(const SetRerankArgs a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= a.n_a) return;
  const float2 *c = a.cand + (static_cast<size_t>(a.dir) * a.rows_pad + row) * a.NB * 4 * 3;
  unsigned long long best[R];
  rerank_row<3, R>(a.a_rows + static_cast<size_t>(row) * 256, a.b_rows, a.n_b, c, a.NB, lane, best);
  if (lane == 0) {
    a.out1[row] = best[0];
    if (R == 2) a.out2[row] = best[R - 1];
  }
}

// mutual nearest neighbour from the two directions' exact keys
__global__ void match_cross_kernel(const unsigned long long *rowbest, const unsigned long long *colbest, int nq, int *q2t, float *dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long rb = rowbest[i];
  int out = -1;
  float d = 0.f;
  if (rb != ~0ull) {
    const int j = static_cast<int>(rb & 0xFFFFFFFFu);
    d = sqrtf(__uint_as_float(static_cast<unsigned>(rb >> 32)));
    if (static_cast<int>(colbest[j] & 0xFFFFFFFFu) == i) out = j;
  }
  q2t[i] = out;
  dist[i] = d;
}

// fp32 rows -> the fp16 copy the tensor-core nomination reads (+ a flag if a row is not a unit vector: the score bound
// above needs |x| = 1, such sets take the CUDA-core exact path); optional gather (rows == nullptr: identity)
__global__ void __launch_bounds__(256) desc_prepare_kernel(const float *__restrict__ src, const int *__restrict__ rows, int n,
                                                           float *__restrict__ dst32, __half *__restrict__ dst16, int *nonunit) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const float *s = src + static_cast<size_t>(rows ? rows[i] : i) * 256 + lane * 8;
  const float4 v0 = *reinterpret_cast<const float4 *>(s), v1 = *reinterpret_cast<const float4 *>(s + 4);
  if (dst32 != nullptr && dst32 + static_cast<size_t>(i) * 256 + lane * 8 != s) {
    *reinterpret_cast<float4 *>(dst32 + static_cast<size_t>(i) * 256 + lane * 8) = v0;
    *reinterpret_cast<float4 *>(dst32 + static_cast<size_t>(i) * 256 + lane * 8 + 4) = v1;
  }
  uint4 o;
  __half2 h;
  h = __floats2half2_rn(v0.x, v0.y); o.x = *reinterpret_cast<uint32_t *>(&h);
  h = __floats2half2_rn(v0.z, v0.w); o.y = *reinterpret_cast<uint32_t *>(&h);
  h = __floats2half2_rn(v1.x, v1.y); o.z = *reinterpret_cast<uint32_t *>(&h);
  h = __floats2half2_rn(v1.z, v1.w); o.w = *reinterpret_cast<uint32_t *>(&h);
  *reinterpret_cast<uint4 *>(dst16 + static_cast<size_t>(i) * 256 + lane * 8) = o;
  float ss = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w + v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
  if (lane == 0 && !(fabsf(ss - 1.0f) <= 1e-3f)) atomicExch(nonunit, 1);
}

}  // namespace spfe
