// tcgen05 implicit-GEMM convolution for the SuperPoint encoder and heads.
//
// Replaces the cuDNN calls behind SPFrontend::forward (reference
// orb_slam2/src/cv/sp_extractor.cpp:82-100): conv1b..conv4b, convPa||convDa
// (3x3, pad 1) and convPb / convDb (1x1), each with its elementwise tail
// (ReLU, 2x2 max-pool, channel L2 norm, or the whole detector head :105-131)
// fused into the TMEM epilogue.
//
// GEMM view per CTA tile:  D[128 pixels, N couts] = sum over (tap, 64-channel
// block) of A[128, 64] * W[N, 64]^T, fp16 operands, fp32 accumulation in TMEM.
//   * activations are NHWC fp16, so a pixel's 64 channels are one 128-byte row of a 128B-swizzled UMMA operand.
//     A work item is HALVES (1 or 2) horizontally adjacent tiles of 8 (x) by 16 (y) output pixels.
//   * im2col is done by TMA + descriptor arithmetic: per item and 64-channel block ONE
//     cp.async.bulk.tensor.4d loads a slab of 18 rows x PW pixels x 64 ch (PW = 8*HALVES + 8, the halo padded to
//     a multiple of 8 pixels) with out-of-bounds zero fill (= the conv's zero padding).  The hardware applies the
//     128B swizzle on absolute shared-memory address bits (verified by tools/umma_shift_test.cu), so the A
//     operand of tap (dy, dx) of half h is simply the same slab addressed at +((dy*PW + h*8 + dx) * 128) bytes
//     with a stride of PW*128 bytes between 8-row groups: every slab byte feeds up to 9 taps, L2 is read
//     1.7x (HALVES = 2) or 2.25x (HALVES = 1) per input element instead of 9x, and a streamed weight block is
//     used by both halves (M = 256 per weight fetch).
//   * weights are packed [tap][cblock][cout][64] fp16; small layers keep all of them resident in shared memory
//     for the CTA's lifetime, large layers stream [N x 64] blocks through a second ring.
//   * persistent CTAs (1 per SM), warp-specialised: warp 0 = slab TMA producer, warp 3 = weight TMA producer,
//     warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue.  Two TMEM accumulator sets so the
//     epilogue of item i overlaps the MMAs of item i+1.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace spfe {

enum { EPI_RELU = 0, EPI_RELU_POOL = 1, EPI_L2NORM = 2, EPI_DETECT = 3, EPI_TOP2 = 4, EPI_TOP3 = 5 };

struct ConvArgs {
  int B, H, W;            // conv spatial size (input == output, before pooling)
  int tiles_x, tiles_y;   // ceil(W/8), ceil(H/16)
  int NB;                 // cout blocks of N channels each (work items per pixel tile)
  int n_items;            // B * tiles_y * tiles_x * NB
  int cin_off;            // first input channel inside the input tensor
  int cin_blk_stride;     // 64-channel blocks between consecutive slabs (0 or 1 = dense; 2 = only the hi blocks of an XP tensor)
  int cout_stride;        // channels per pixel of the output tensor
  const float *bias;      // [NB*N]
  __half *out;            // RELU / RELU_POOL / L2NORM: NHWC fp16
  // EPI_DETECT outputs (one value per 8x8 cell unless noted)
  float *score;           // max softmax prob over the 64 position channels
  uint8_t *argmax;        // its channel index
  float *semi_dust;       // raw dustbin logit
  float *dense_dust;      // dustbin softmax prob
  float *heat_log;        // [B][8H][8W] log(clamp(p, 1e-3)) depth-to-space, or nullptr
  unsigned *heat_minmax;  // [B][2] ordered-uint min / max of heat_log, or nullptr
  // EPI_TOP2 (descriptor matching as a GEMM, see match.cuh): "frame slot" s holds descriptor rows of one frame;
  // item = (pair z, direction, 128-row tile, 256-column block): A = slot z+1-dir, B = slot z+dir
  const int *m_count;     // [slots] valid rows per slot
  float2 *m_cand;         // [2][Z][rows_pad][NB][4][TOPK] (score, index-as-float-bits): best TOPK per row, 256-column block
                          // and column residue mod 4 (a "virtual block" of 64 columns)
  int m_rows_pad;         // rows per slot in the fp16 descriptor tensor (multiple of 256)
  int m_tiles;            // 128-row tiles per slot that can hold valid rows
  // descriptor-SET mode (spfe_match_*: one direction per launch): tmA = the A set's rows, tmW = the B set's rows, both
  // tensors hold a single slot; row counts are host-known
  int m_set, m_na, m_nb, m_dir;
};

template <int TAPS_, int CB_, int N_, int EPI_, bool WRES_, int SA_, int SB_, int HALVES_ = 1, int EG_ = 1, bool PAIR_ = false, bool XP_ = false>
struct ConvCfg {
  // XP ("exact" mode, SPFE_EXACT): fp32-equivalent products out of fp16 tensor-core operands.  Activations and weights
  // are held as hi + lo fp16 pairs (22 significant bits): an activation tensor has its channels interleaved in 64-blocks
  // [hi_0 | lo_0 | hi_1 | lo_1 ...], the packed weights likewise [Wh_0 | Wl_0 | Wh_1 | Wl_1 ...] per tap, and a product is
  // the three MMAs  Ah*Wh + Ah*Wl + Al*Wh  (Al*Wl, 2^-24 relative, is dropped) accumulated in the same fp32 TMEM
  // accumulator.  CB counts SLABS (= 64-channel blocks of the XP tensor = 2 x real blocks): an even slab (hi) is
  // multiplied with two weight sets (Wh, Wl), an odd one (lo) with Wh only.  RELU / RELU_POOL epilogues write hi and lo.
  static constexpr bool XP = XP_;
  static_assert(!XP_ || CB_ % 2 == 0, "XP: slabs come in (hi, lo) pairs");
  // PAIR: two CTAs of a cluster run every MMA together (cta_group::2, M = 256): each works on its own item (its own
  // slabs, accumulators and epilogue) but holds only half of the weights, so the shared-memory operand reads per SM and
  // MMA drop from A + B to A + B/2 -- the limiter of the N = 64 layers (tools/umma2_rate.cu).  The leader (rank 0)
  // issues the MMAs for both; items 2q and 2q+1 go to ranks 0 and 1 of pair q mod (grid / 2).
  static constexpr bool PAIR = PAIR_;
  static_assert(!PAIR_ || (N_ % 32 == 0 && EPI_ != EPI_TOP2 && EPI_ != EPI_TOP3), "pairs: convolution layers with N a multiple of 32");
  // EG = epilogue warp groups (4 warps each).  With 2, group g drains accumulator set g, so the epilogues of two
  // consecutive items run side by side: for the layers whose epilogue (softmax / log / norm), not the MMAs, paces the CTA.
  static constexpr int EG = EG_, THREADS = 128 + 128 * EG_;
  static_assert(EG_ == 1 || EG_ == 2, "one or two epilogue groups");
  static constexpr bool MATCH = (EPI_ == EPI_TOP2 || EPI_ == EPI_TOP3);
  static constexpr int TOPK = EPI_ == EPI_TOP3 ? 3 : 2;  // candidates nominated per row and 256-column block
  static constexpr int TAPS = TAPS_, CB = CB_, N = N_, EPI = EPI_, SA = SA_, SB = SB_, HALVES = HALVES_;
  static constexpr bool WRES = WRES_;
  static constexpr int NDX = TAPS == 9 ? 3 : 1;
  static constexpr int NDY = NDX;
  static constexpr int TILE_W = 8 * HALVES;                       // output pixels per item row
  static constexpr int PW = TAPS == 9 ? 8 * HALVES + 8 : 8;       // slab pixels per image row (halo padded to x8)
  static constexpr int SLAB_ROWS = TAPS == 9 ? 18 : 16;
  static constexpr int SLAB_BYTES = SLAB_ROWS * PW * 128;
  static constexpr int SBO = PW * 128;                            // bytes between 8-row groups of the A operand
  static constexpr int BBLK_BYTES = (PAIR_ ? N_ / 2 : N_) * 128;  // a pair member holds N/2 rows of every weight block
  static constexpr int NWB = TAPS * CB;
  static constexpr int B_BYTES = (WRES ? NWB : SB) * BBLK_BYTES;
  static constexpr int ACC_STRIDE = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = 2 * HALVES * ACC_STRIDE;       // two accumulator sets of HALVES tiles
  static constexpr int NBAR = 2 * SA + 2 * SB + 5;
  // dynamic shared memory: [1024 align slack][A ring][B ring / resident W][barriers][tmem slot][bias NB*N f32]
  static constexpr int SMEM_FIXED = 1024 + SA * SLAB_BYTES + B_BYTES + NBAR * 8 + 16;
  static constexpr int smem_bytes(int nb) { return SMEM_FIXED + nb * N * 4; }
  static_assert(N % 16 == 0 && N <= 256, "UMMA M=128 needs N % 16 == 0, N <= 256");
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation: power of two <= 512 columns");
  static_assert(HALVES == 1 || TAPS == 9, "tile pairs only for 3x3 convolutions");
  static_assert(SMEM_FIXED + N * 4 <= 232448, "shared memory budget");
};

__device__ __forceinline__ unsigned f32_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float f32_from_ordered(unsigned u) {
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}

// hi / lo split of two fp32 values: hi = fp16(v), lo = fp16(v - hi)  (XP activations)
__device__ __forceinline__ void split_h2(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}
// channel c of an XP tensor: hi at (c / 64) * 128 + c % 64, lo 64 further
__device__ __forceinline__ int xp_off(int c) { return ((c >> 6) << 7) + (c & 63); }

// Epilogue: bias + ReLU + 2x2 max-pool of one 8x16-pixel accumulator tile (lane = h*8 + w), fp16 NHWC store.
// The pool partners are lane^1 (x) and lane^8 (y), both inside the warp.  The 2x2 maximum is a reduce-scatter: every
// 32-channel chunk is first halved with the x partner (each lane keeps the 16 channels whose bit 3 equals its x parity
// and receives the partner's values for them), then halved again with the y partner (bit 4 / y parity), so each of the
// four lanes of a 2x2 group ends up with the 8-channel slice it stores -- 24 shuffles per chunk instead of 64 (this
// epilogue, not the MMAs, paced the pooled 64 -> 64 layers).  max is exact, so the result does not depend on the order.
// All TMEM loads of the tile are issued before the first wait.
template <int N, bool XP = false>
__device__ __forceinline__ void epilogue_relu_pool(uint32_t taddr, const float *bias, int lane, int hl, int wl, int x0,
                                                   int y0, int b, int nb, int H, int W, int cout_stride, __half *out) {
  static_assert(N == 64 || N == 128, "pooled layers have 64 or 128 output channels");
  const int Ho = H >> 1, Wo = W >> 1;
  const int yo = (y0 >> 1) + (hl >> 1), xo = (x0 >> 1) + (wl >> 1);
  const bool valid = (yo < Ho) && (xo < Wo);
  const bool px = lane & 1, py = (lane >> 3) & 1;
  const int q = (lane & 1) | (((lane >> 3) & 1) << 1);
  __half *dst = out + ((static_cast<size_t>(b) * Ho + yo) * Wo + xo) * cout_stride + (XP ? 0 : nb * N + q * 8);
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 64) {
    float v[64];
    tmem_ld16(taddr + c0, v);
    tmem_ld16(taddr + c0 + 16, v + 16);
    tmem_ld16(taddr + c0 + 32, v + 32);
    tmem_ld16(taddr + c0 + 48, v + 48);
    tmem_ld_wait();
#pragma unroll
    for (int ch = 0; ch < 64; ch += 32) {
      float u[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {  // x partner: channels (j >> 3) * 16 + px * 8 + (j & 7) of this chunk
        const int ca = ch + (j >> 3) * 16 + (j & 7), cb = ca + 8;
        const float keep = px ? v[cb] : v[ca], send = px ? v[ca] : v[cb];
        u[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
      }
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {   // y partner: channels py * 16 + px * 8 + j == q * 8 + j
        const float keep = py ? u[8 + j] : u[j], send = py ? u[j] : u[8 + j];
        w[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
        w[j] = fmaxf(w[j] + bias[c0 + ch + q * 8 + j], 0.f);
      }
      if (valid) {
        if constexpr (XP) {
          uint4 o, l;
          split_h2(w[0], w[1], o.x, l.x);
          split_h2(w[2], w[3], o.y, l.y);
          split_h2(w[4], w[5], o.z, l.z);
          split_h2(w[6], w[7], o.w, l.w);
          __half *d = dst + xp_off(nb * N + c0 + ch + q * 8);
          *reinterpret_cast<uint4 *>(d) = o;
          *reinterpret_cast<uint4 *>(d + 64) = l;
        } else {
        uint4 o;
        o.x = pack_h2(w[0], w[1]);
        o.y = pack_h2(w[2], w[3]);
        o.z = pack_h2(w[4], w[5]);
        o.w = pack_h2(w[6], w[7]);
        *reinterpret_cast<uint4 *>(dst + c0 + ch) = o;
        }
      }
    }
  }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const ConvArgs p) {
  constexpr int N = Cfg::N, CB = Cfg::CB, SA = Cfg::SA, SB = Cfg::SB;
  const int NB = p.NB;  // resident weights (WRES) require NB == 1 (checked on the host)
  constexpr int NDX = Cfg::NDX, NDY = Cfg::NDY, EPI = Cfg::EPI;
  constexpr bool WRES = Cfg::WRES, XP = Cfg::XP;
  // XP: weight sets per slab (hi slabs meet Wh and Wl, lo slabs Wh only) and the packed weight block of (slab, set)
  auto n_wsets = [](int cb) { return (XP && !(cb & 1)) ? 2 : 1; };
  auto w_block = [](int cb, int ws) { return XP ? ((cb & ~1) + ((cb & 1) ? 0 : ws)) : cb; };
  const int cin_step = 64 * (p.cin_blk_stride > 1 ? p.cin_blk_stride : 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sA = smem;
  uint8_t *sB = sA + SA * Cfg::SLAB_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + Cfg::B_BYTES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + Cfg::NBAR);
  float *sBias = reinterpret_cast<float *>(tmem_slot + 4);

  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * SA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * SA + SB + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * SA + 2 * SB + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * SA + 2 * SB + 2 + s); };
  const uint32_t w_full = bar0 + 8u * (2 * SA + 2 * SB + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < SB; s++) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; s++) { mbar_init(t_full(s), 1); mbar_init(t_empty(s), Cfg::PAIR ? 8 : 128); }
    mbar_init(w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if constexpr (Cfg::PAIR) cluster_sync_all();  // both CTAs' barriers exist before anything can signal across the pair
  if (warp == 2) {
    if constexpr (Cfg::PAIR) { tmem_alloc_pair(smem_u32(tmem_slot), Cfg::TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  if constexpr (!Cfg::MATCH)
    for (int i = threadIdx.x; i < NB * N; i += blockDim.x) sBias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, bias) and the resident-weight loads below do not
  // depend on the previous kernel of the stream, so this CTA may have started while that kernel was still draining;
  // the roles that touch its output (activation loads, global stores) wait for it first.
  griddep_launch_dependents();
  const uint32_t rank = Cfg::PAIR ? cluster_ctarank() : 0u;
  // Work distribution.  Plain: CTA b takes items b, b + grid, ...  Pair: pair j = b / 2 takes item pairs j, j + grid / 2, ...;
  // rank r works on tile 2q + r (clamped: with an odd tile count the last tile is computed by both ranks, which write
  // identical results).
  // With several cout blocks (NB > 1) the two ranks take the same block nb of two neighbouring tiles, so that they share B.
  const int q_first = Cfg::PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int q_step = Cfg::PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int n_tiles = p.n_items / NB;
  const int q_end = Cfg::PAIR ? ((n_tiles + 1) >> 1) * NB : p.n_items;
  auto item_of = [&](int q) {
    if constexpr (!Cfg::PAIR) return q;
    const int tq = q / NB, nb = q - tq * NB;
    return min(2 * tq + static_cast<int>(rank), n_tiles - 1) * NB + nb;
  };
  if constexpr (Cfg::PAIR && WRES) {  // each rank loads its half of the resident weights; nobody starts before both halves are in
    if (warp == 3) {
      if (lane == 0) {
        mbar_expect_tx(w_full, Cfg::NWB * Cfg::BBLK_BYTES);
        for (int wb = 0; wb < Cfg::NWB; wb++)
          tma_load_2d(smem_u32(sB + wb * Cfg::BBLK_BYTES), &tmW, w_full, 0, wb * N + static_cast<int>(rank) * (N / 2));
      }
      mbar_wait(w_full, 0);
    }
    cluster_sync_all();
  }

  auto decode = [&](int item, int &nb, int &x0, int &y0, int &b) {
    nb = item % NB;
    int t = item / NB;
    if constexpr (Cfg::MATCH) {  // rows of slot b as an [rows/8][8] "image": tile mt covers rows mt*128 .. +127
      x0 = 0;
      y0 = (t % p.m_tiles) * 16;
      t /= p.m_tiles;
      b = p.m_set ? 0 : (t >> 1) + 1 - (t & 1);  // pair z = t >> 1, direction = t & 1
      return;
    }
    x0 = (t % p.tiles_x) * Cfg::TILE_W;
    t /= p.tiles_x;
    y0 = (t % p.tiles_y) * 16;
    b = t / p.tiles_y;
  };

  if (warp == 0) {
    // ------------------------------------------------ activation slab producer
    if (lane == 0) {
      griddep_wait();
      uint32_t it = 0;
      for (int q = q_first; q < q_end; q += q_step) {
        int nb, x0, y0, b;
        decode(item_of(q), nb, x0, y0, b);
        for (int cb = 0; cb < CB; cb++, it++) {
          const int s = it % SA;
          mbar_wait(a_empty(s), ((it / SA) & 1) ^ 1);
          if constexpr (Cfg::PAIR) {  // both ranks' slabs complete on the leader's barrier (the leader issues the MMAs)
            if (rank == 0) mbar_expect_tx(a_full(s), 2 * Cfg::SLAB_BYTES);
            tma_load_4d_pair(smem_u32(sA + s * Cfg::SLAB_BYTES), &tmA, mapa_shared(a_full(s), 0), p.cin_off + cb * cin_step,
                             x0 - (NDX == 3 ? 1 : 0), y0 - (NDY == 3 ? 1 : 0), b);
          } else {
            mbar_expect_tx(a_full(s), Cfg::SLAB_BYTES);
            tma_load_4d(smem_u32(sA + s * Cfg::SLAB_BYTES), &tmA, a_full(s), p.cin_off + cb * cin_step,
                        x0 - (NDX == 3 ? 1 : 0), y0 - (NDY == 3 ? 1 : 0), b);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ weight producer
    if (lane == 0 && !(Cfg::PAIR && WRES)) {
      if (WRES) {
        mbar_expect_tx(w_full, Cfg::NWB * Cfg::BBLK_BYTES);
        for (int wb = 0; wb < Cfg::NWB; wb++)
          tma_load_2d(smem_u32(sB + wb * Cfg::BBLK_BYTES), &tmW, w_full, 0, wb * N);
      } else {
        if constexpr (Cfg::MATCH) griddep_wait();  // the B operand is the previous kernel's output here, not constant weights
        uint32_t jt = 0;
        for (int q = q_first; q < q_end; q += q_step) {
          const int item = item_of(q);
          const int nb = item % NB;  // (a pair's two items share nb: NB divides 2 or is 1, see the host-side check)
          for (int cb = 0; cb < CB; cb++)
           for (int ws = 0; ws < n_wsets(cb); ws++)
            for (int dx = 0; dx < NDX; dx++)
              for (int dy = 0; dy < NDY; dy++, jt++) {
                const int s = jt % SB;
                mbar_wait(b_empty(s), ((jt / SB) & 1) ^ 1);
                const int wb = (dy * NDX + dx) * CB + w_block(cb, ws);
                if constexpr (Cfg::PAIR) {  // each rank streams its N/2 rows of the block; both halves complete on the leader's barrier
                  if (rank == 0) mbar_expect_tx(b_full(s), 2 * Cfg::BBLK_BYTES);
                  tma_load_2d_pair(smem_u32(sB + s * Cfg::BBLK_BYTES), &tmW, mapa_shared(b_full(s), 0), 0,
                                   (wb * NB + nb) * N + static_cast<int>(rank) * (N / 2));
                  continue;
                }
                mbar_expect_tx(b_full(s), Cfg::BBLK_BYTES);
                if constexpr (Cfg::MATCH) {
                  const int t = item / NB / p.m_tiles;  // B operand = descriptor rows of the other slot of the pair
                  const int bslot = p.m_set ? 0 : (t >> 1) + (t & 1);
                  tma_load_2d(smem_u32(sB + s * Cfg::BBLK_BYTES), &tmW, b_full(s), cb * 64, bslot * p.m_rows_pad + nb * N);
                } else {
                  tma_load_2d(smem_u32(sB + s * Cfg::BBLK_BYTES), &tmW, b_full(s), 0, (wb * NB + nb) * N);
                }
              }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    // The whole warp runs the loop so that every address / descriptor stays warp-uniform (uniform datapath); one
    // elected lane issues.  Entering an elected region costs ~90 cycles (measured, tools/umma_rate.cu) while an
    // M128 x N<=128 MMA needs only 50-64, so MMAs are issued in the largest groups the buffering allows: a whole
    // item when the weights are resident, one horizontal tap (3 dy x HALVES x 4 k-steps) when they are streamed.
    constexpr uint32_t idesc = Cfg::PAIR ? umma_idesc_f16_pair(N) : umma_idesc_f16(N);
    constexpr int PW = Cfg::PW, HALVES = Cfg::HALVES;
    uint32_t it = 0, jt = 0, tcount = 0;
    if (WRES) mbar_wait(w_full, 0);
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    // A operand of tap (dy, dx), half h, k-step k: slab + ((dy*PW + h*8 + dx)*128 + k*32) bytes, in (addr >> 4) units
    auto a_off = [](int dy, int dx, int h, int k) { return static_cast<uint64_t>((dy * PW + h * 8 + dx) * 8 + 2 * k); };
    for (int q = q_first; q < q_end && rank == 0; q += q_step, tcount++) {  // in a pair only the leader issues
      const uint32_t acc = tcount & 1;
      mbar_wait(t_empty(acc), ((tcount >> 1) & 1) ^ 1);
      const uint32_t d_tmem = tmem_base + acc * HALVES * Cfg::ACC_STRIDE;
      if constexpr (WRES) {
        static_assert(!WRES || SA >= CB, "resident-weight kernels issue a whole item at once: all its slabs must fit the ring");
        uint32_t st[CB];
#pragma unroll
        for (int cb = 0; cb < CB; cb++) {
          st[cb] = (it + cb) % SA;
          mbar_wait(a_full(st[cb]), ((it + cb) / SA) & 1);
        }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int cb = 0; cb < CB; cb++) {
            const uint64_t a0 = umma_desc_sw128(sA_u + st[cb] * Cfg::SLAB_BYTES, Cfg::SBO);
#pragma unroll
           for (int ws = 0; ws < ((XP && !(cb & 1)) ? 2 : 1); ws++)
#pragma unroll
            for (int dx = 0; dx < NDX; dx++)
#pragma unroll
              for (int dy = 0; dy < NDY; dy++) {
                const uint64_t b0 = umma_desc_sw128(sB_u + ((dy * NDX + dx) * CB + w_block(cb, ws)) * Cfg::BBLK_BYTES, 1024);
#pragma unroll
                for (int h = 0; h < HALVES; h++)
#pragma unroll
                  for (int k = 0; k < 4; k++) {
                    if constexpr (Cfg::PAIR)
                      umma_f16_pair(d_tmem + h * Cfg::ACC_STRIDE, a0 + a_off(dy, dx, h, k), b0 + 2 * k, idesc, (cb | ws | dx | dy | k) ? 1u : 0u);
                    else
                      umma_f16(d_tmem + h * Cfg::ACC_STRIDE, a0 + a_off(dy, dx, h, k), b0 + 2 * k, idesc, (cb | ws | dx | dy | k) ? 1u : 0u);
                  }
              }
            // the slab is free as soon as its own MMAs retire (keeps the TMA ring busy); in a pair both ranks are told
            if constexpr (Cfg::PAIR) umma_commit_pair(a_empty(st[cb])); else umma_commit(a_empty(st[cb]));
          }
          if constexpr (Cfg::PAIR) umma_commit_pair(t_full(acc)); else umma_commit(t_full(acc));
        }
        __syncwarp();
        it += CB;
      } else {
        static_assert(WRES || SB >= NDY, "streamed weights: one horizontal tap's weight blocks must fit the ring");
        // group = GDY vertical taps per elected region: all three when an MMA is short (N <= 128), one when N = 256
        // (128 cycles per MMA dwarf the region overhead and the weight ring is only 4 blocks deep)
        constexpr int GDY = (N >= 256 || SB < 2 * NDY) ? 1 : NDY;
#pragma unroll 1
        for (int cb = 0; cb < CB; cb++, it++) {
          const uint32_t s = it % SA;
          mbar_wait(a_full(s), (it / SA) & 1);
          const uint64_t a0 = umma_desc_sw128(sA_u + s * Cfg::SLAB_BYTES, Cfg::SBO);
          const int nws = n_wsets(cb);
#pragma unroll 1
         for (int ws = 0; ws < nws; ws++) {
#pragma unroll
          for (int dx = 0; dx < NDX; dx++) {
#pragma unroll
            for (int g = 0; g < NDY; g += GDY) {
              uint32_t sb[GDY];
#pragma unroll
              for (int d = 0; d < GDY; d++) {
                sb[d] = (jt + d) % SB;
                mbar_wait(b_full(sb[d]), ((jt + d) / SB) & 1);
              }
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int d = 0; d < GDY; d++) {
                  const int dy = g + d;
                  const uint64_t b0 = umma_desc_sw128(sB_u + sb[d] * Cfg::BBLK_BYTES, 1024);
#pragma unroll
                  for (int h = 0; h < HALVES; h++)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                      if constexpr (Cfg::PAIR)
                        umma_f16_pair(d_tmem + h * Cfg::ACC_STRIDE, a0 + a_off(dy, dx, h, k), b0 + 2 * k, idesc, (cb | ws | dx | dy | k) ? 1u : 0u);
                      else
                        umma_f16(d_tmem + h * Cfg::ACC_STRIDE, a0 + a_off(dy, dx, h, k), b0 + 2 * k, idesc, (cb | ws | dx | dy | k) ? 1u : 0u);
                    }
                  if constexpr (Cfg::PAIR) umma_commit_pair(b_empty(sb[d])); else umma_commit(b_empty(sb[d]));
                }
                if (dx == NDX - 1 && g + GDY >= NDY && ws == nws - 1) {
                  if constexpr (Cfg::PAIR) umma_commit_pair(a_empty(s)); else umma_commit(a_empty(s));
                  if (cb == CB - 1) { if constexpr (Cfg::PAIR) umma_commit_pair(t_full(acc)); else umma_commit(t_full(acc)); }
                }
              }
              __syncwarp();
              jt += GDY;
            }
          }
         }  // weight sets
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (TMEM -> regs -> global)
    const int wq = warp & 3;
    const int hl = wq * 4 + (lane >> 3), wl = lane & 7;
    const uint32_t egroup = static_cast<uint32_t>(warp - 4) >> 2;
    griddep_wait();
    uint32_t tcount = 0;
    for (int q = q_first; q < q_end; q += q_step, tcount++) {
      if (Cfg::EG == 2 && (tcount & 1) != egroup) continue;  // the other group's accumulator set
      const int item = item_of(q);
      int nb, x0_item, y0, b;
      decode(item, nb, x0_item, y0, b);
      const int acc = tcount & 1;
      mbar_wait(t_full(acc), (tcount >> 1) & 1);
      tc_fence_after();
      const float *bias = sBias + nb * N;
#pragma unroll 1
      for (int half = 0; half < Cfg::HALVES; half++) {  // the item's 8-pixel-wide tiles, one accumulator each
      const int x0 = x0_item + half * 8;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + (acc * Cfg::HALVES + half) * Cfg::ACC_STRIDE;
      const int y = y0 + hl, x = x0 + wl;

      if constexpr (EPI == EPI_RELU) {
        const bool valid = (y < p.H) && (x < p.W);
        __half *dst = p.out + ((static_cast<size_t>(b) * p.H + y) * p.W + x) * p.cout_stride + (XP ? 0 : nb * N);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld16(taddr + c0, v);
          tmem_ld16(taddr + c0 + 16, v + 16);
          tmem_ld_wait();
          if (valid) {  // 64 bytes of this pixel: two 32-byte stores (full sectors)
#pragma unroll
            for (int g = 0; g < 4; g += 2) {
              uint4 o[2];
              uint32_t *ow = reinterpret_cast<uint32_t *>(o);
              if constexpr (XP) {  // hi and lo halves of the same 16 channels, 64 channels apart
                uint4 l[2];
                uint32_t *lw = reinterpret_cast<uint32_t *>(l);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  const int c = g * 8 + j * 2;
                  split_h2(fmaxf(v[c] + bias[c0 + c], 0.f), fmaxf(v[c + 1] + bias[c0 + c + 1], 0.f), ow[j], lw[j]);
                }
                __half *d = dst + xp_off(nb * N + c0 + g * 8);
                st_global_v8(d, o[0], o[1]);
                st_global_v8(d + 64, l[0], l[1]);
              } else {
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const int c = g * 8 + j * 2;
                ow[j] = pack_h2(fmaxf(v[c] + bias[c0 + c], 0.f), fmaxf(v[c + 1] + bias[c0 + c + 1], 0.f));
              }
              st_global_v8(dst + c0 + g * 8, o[0], o[1]);
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_RELU_POOL) {
        epilogue_relu_pool<N, XP>(taddr, bias, lane, hl, wl, x0, y0, b, nb, p.H, p.W, p.cout_stride, p.out);
      } else if constexpr (Cfg::MATCH) {
        // One thread = one descriptor row of the A slot: best TOPK dot products (first index on ties) among this
        // item's 256 columns of the B slot.  The exact fp32 re-rank happens in match.cuh.
        constexpr int TOPK = Cfg::TOPK;
        const int t = item / NB / p.m_tiles;
        const int z = p.m_set ? 0 : t >> 1, dir = p.m_set ? p.m_dir : t & 1;
        const int n_a = p.m_set ? p.m_na : p.m_count[b], n_b = p.m_set ? p.m_nb : p.m_count[z + dir];
        const int row = y0 * 8 + wq * 32 + lane;
        // Sortable 32-bit keys: ordered score bits with the low byte replaced by (255 - column) -- a larger key is a
        // larger score, ties go to the lower column.  Dropping 8 mantissa bits (2^-15 relative) is harmless: the keys
        // only nominate candidates.  Four independent (best, second[, third]) chains keep the dependency depth short.
        unsigned m1[4] = {0u, 0u, 0u, 0u}, m2[4] = {0u, 0u, 0u, 0u}, m3[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld16(taddr + c0, v);
          tmem_ld16(taddr + c0 + 16, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int cl = c0 + j;
            unsigned key = (f32_ordered(v[j]) & 0xFFFFFF00u) | static_cast<unsigned>(255 - cl);
            key = (nb * N + cl < n_b) ? key : 0u;
            const unsigned lo = min(m1[j & 3], key);
            m1[j & 3] = max(m1[j & 3], key);
            if constexpr (TOPK == 3) {
              const unsigned lo2 = min(m2[j & 3], lo);
              m2[j & 3] = max(m2[j & 3], lo);
              m3[j & 3] = max(m3[j & 3], lo2);
            } else {
              m2[j & 3] = max(m2[j & 3], lo);
            }
          }
        }
        // The four chains are NOT merged: chain k holds the best TOPK of the 64 columns == k (mod 4), a "virtual block".
        // Near-ties then rarely share a block (the exact re-scan of match.cuh triggers 16x less often than with one list
        // per 256 columns) and a re-scan touches 64 columns instead of 256.
        if (row < n_a) {
          float2 *dst = p.m_cand + (((static_cast<size_t>(dir) * p.B + z) * p.m_rows_pad + row) * NB + nb) * 4 * TOPK;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const unsigned bk[3] = {m1[k], m2[k], m3[k]};
#pragma unroll
            for (int e = 0; e < TOPK; e++)
              dst[k * TOPK + e] = make_float2(bk[e] ? f32_from_ordered(bk[e] & 0xFFFFFF00u) : -INFINITY,
                                              __int_as_float(bk[e] ? nb * N + 255 - static_cast<int>(bk[e] & 0xFFu) : -1));
          }
        }
      } else if constexpr (EPI == EPI_L2NORM) {
        // convDb + channel-wise L2 normalisation (sp_extractor.cpp:100-103); N == all 256 channels.
        const bool valid = (y < p.H) && (x < p.W);
        __half *dst = p.out + ((static_cast<size_t>(b) * p.H + y) * p.W + x) * p.cout_stride + nb * N;
        float ss = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld16(taddr + c0, v);
          tmem_ld16(taddr + c0 + 16, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const float t = v[j] + bias[c0 + j];
            ss = fmaf(t, t, ss);
          }
        }
        const float inv = 1.0f / sqrtf(ss);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          float v[32];
          tmem_ld16(taddr + c0, v);
          tmem_ld16(taddr + c0 + 16, v + 16);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int g = 0; g < 4; g += 2) {
              uint4 o[2];
              uint32_t *ow = reinterpret_cast<uint32_t *>(o);
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const int c = g * 8 + j * 2;
                ow[j] = pack_h2((v[c] + bias[c0 + c]) * inv, (v[c + 1] + bias[c0 + c + 1]) * inv);
              }
              st_global_v8(dst + c0 + g * 8, o[0], o[1]);
            }
          }
        }
      } else {
        // EPI_DETECT: convPb logits (65 of the 80 columns) -> detector head, all thread-local:
        // softmax over 65 (:105), dustbin logit/prob (:106-107), max/argmax over 64 (:112-114),
        // log(clamp(p, 1e-3)) depth-to-space heat (:129-131).  One thread == one 8x8 cell.
        static_assert(EPI != EPI_DETECT || N == 80, "detector head expects N = 80 (65 padded)");
        const bool valid = (y < p.H) && (x < p.W);
        float l[80];
#pragma unroll
        for (int c0 = 0; c0 < 80; c0 += 16) tmem_ld16(taddr + c0, l + c0);
        tmem_ld_wait();
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 65; j++) {
          l[j] += bias[j];
          m = fmaxf(m, l[j]);
        }
        const float dust_logit = l[64];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 65; j++) {
          l[j] = expf(l[j] - m);
          sum += l[j];
        }
        float best = -1.f;
        int arg = 0;
#pragma unroll
        for (int j = 0; j < 64; j++) {
          l[j] = l[j] / sum;
          if (l[j] > best) { best = l[j]; arg = j; }
        }
        const float dust_p = l[64] / sum;
        float hmin = INFINITY, hmax = -INFINITY;
        if (valid) {
          const size_t cell = (static_cast<size_t>(b) * p.H + y) * p.W + x;
          p.score[cell] = best;
          p.argmax[cell] = static_cast<uint8_t>(arg);
          p.semi_dust[cell] = dust_logit;
          p.dense_dust[cell] = dust_p;
          if (p.heat_log != nullptr) {
            const int Wf = p.W * 8;
            float *hrow = p.heat_log + (static_cast<size_t>(b) * p.H * 8 + static_cast<size_t>(y) * 8) * Wf + x * 8;
#pragma unroll
            for (int r = 0; r < 8; r++) {
              float h[8];
#pragma unroll
              for (int c = 0; c < 8; c++) {
                h[c] = logf(fmaxf(l[r * 8 + c], 0.001f));
                hmin = fminf(hmin, h[c]);
                hmax = fmaxf(hmax, h[c]);
              }
              st_global_v8(hrow + static_cast<size_t>(r) * Wf,  // the cell's 8 pixels of this row: one 32-byte store
                           make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(h[2]), __float_as_uint(h[3])),
                           make_uint4(__float_as_uint(h[4]), __float_as_uint(h[5]), __float_as_uint(h[6]), __float_as_uint(h[7])));
            }
          }
        }
        if (p.heat_log != nullptr && p.heat_minmax != nullptr) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            hmin = fminf(hmin, __shfl_xor_sync(0xffffffffu, hmin, o));
            hmax = fmaxf(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
          }
          if (lane == 0 && hmin <= hmax) {
            atomicMin(p.heat_minmax + 2 * b, f32_ordered(hmin));
            atomicMax(p.heat_minmax + 2 * b + 1, f32_ordered(hmax));
          }
        }
      }
      }  // half
      tc_fence_before();
      if constexpr (Cfg::PAIR) {  // the leader's MMA warp waits for both epilogues: one cluster-scope arrival per warp
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(t_empty(acc), 0));
      } else {
        mbar_arrive(t_empty(acc));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (Cfg::PAIR) cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (Cfg::PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace spfe
