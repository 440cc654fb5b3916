// Fused conv1a + conv1b + ReLU + 2x2 max-pool with BOTH layers on the tensor core
// (reference sp_extractor.cpp:81-83, fed by :386-390).
//
// conv1ab.cuh computes conv1a on the CUDA cores and is paced by those FFMA chains (~6000 cycles per item against
// ~4300 for conv1b's 72 MMAs).  Here conv1a is a tcgen05 GEMM as well:
//
//   D1[pixel, 64 ch] = A1[pixel, 16] * W1[64, 16]^T          (K = 9 taps + 1 bias column, padded to 16)
//
//   * A1 (im2col of the 18 x 18 halo of a 16 x 16 output item, 324 rows padded to 3 x 128) is built by four producer
//     warps from the 20 x 20 u8 image patch: pixel values 0..255 are exact in fp16; column 9 is 255 (the bias
//     column); a halo pixel outside the image gets an all-zero row, so its conv1a output is relu(0) = 0 == conv1b's
//     zero padding.
//   * conv1a is the precision-critical layer (|w| up to 197, fp32 in the reference): the fp32 weights are split on
//     the host into hi + lo fp16 parts (22 significant bits) and each M-tile is two K=16 MMAs, A1*W1hi^T + A1*W1lo^T,
//     accumulated in fp32 in TMEM.  The scale 1/255 (cv::Mat::convertTo, :386) is applied in fp32 in the epilogue.
//   * A1 / W1 use the un-swizzled K-major canonical layout (8-row x 16-byte core matrices, LBO = 128 B between the
//     two K chunks, SBO = 256 B between 8-row groups).
//   * four "conv1a epilogue" warps read D1 from TMEM, scale, ReLU, convert to fp16 and store every halo pixel once
//     into the 128B-swizzled slab that conv1b's 72 MMAs read (same slab geometry as conv1ab.cuh / conv_tc.cuh).
//
// Warps: 1 = MMA issuer; 0, 2, 3 = im2col producers (warp 0 also issues the weight loads, warp 2 owns the TMEM
// allocation); 4-7 = conv1b epilogue; 8-11 / 12-15 = conv1a epilogue, channels 0-31 / 32-63.  The issuer interleaves
// c1a(i+1), c1b(i), c1a(i+2), c1b(i+1) ...  so the conv1a epilogue of item i+1 runs under the 72 MMAs of item i and the
// tensor pipe only idles for the six small conv1a MMAs.  The chain  conv1a MMAs -> TMEM -> slab -> barrier -> next conv1a
// MMAs  must fit under one conv1b MMA group; it is the reason for the two epilogue groups and the single e1_done barrier.
//
// PAIR: two CTAs of a cluster run every MMA together (cta_group::2, see conv_tc.cuh / ptx.cuh): each rank keeps its own
// item (patch, im2col, slab, accumulators, epilogues) but only half of the conv1b and conv1a weights (32 couts); the
// leader issues all MMAs, so every "ready" barrier lives in the leader and is armed by both ranks (one arrival per
// warp), while the "done" barriers are signalled in both CTAs by multicast commits.
#pragma once
#include "conv1ab.cuh"

namespace spfe {

#ifdef SPFE_C1M_TRACE
// [n][0] MMA warp: c1a(n) issued, [1] warp 8: c_full(n) seen, [2] s_empty seen, [3] work done, [4] arrived,
// [5] MMA warp: e1_done(n) seen, [6] c1b(n) issued; [7] peer CTA warp 8: work duration, [8] peer: c_full wait, [9] peer arrive duration
__device__ long long g_c1m_trace[16][10];
#define C1M_T(n, k) do { if ((n) >= 200 && (n) < 216 && blockIdx.x == 0 && lane == 0) g_c1m_trace[(n) - 200][k] = clock64(); } while (0)
#else
#define C1M_T(n, k) do { } while (0)
#endif

namespace c1m {
constexpr int PW = c1ab::PW, HALO = c1ab::HALO, STAGE = c1ab::STAGE, WBLK = c1ab::WBLK, WBYTES = c1ab::WBYTES;
constexpr int PATCH = c1ab::PATCH;
constexpr int NPIX = HALO * HALO;          // 324 halo pixels == im2col rows
constexpr int MT = 3;                      // M-tiles of 128 rows
constexpr int A1_BYTES = MT * 128 * 32;    // one im2col stage: 384 rows x 16 fp16
constexpr int W1_PART = 64 * 32;           // one fp16 part (hi or lo) of the conv1a weights [64][16]
constexpr int W1_BYTES = 2 * W1_PART;
constexpr int THREADS = 512, NPROD = 96;   // im2col producers: warps 0, 2, 3
constexpr int NBAR = 14;
constexpr int TMEM_COLS = 512;             // conv1b: 2 stages x 2 halves x 64; conv1a: 3 x 64 at column 256
constexpr int C1A_COL = 256;
constexpr int SMEM = 1024 + 2 * STAGE + WBYTES + 2 * A1_BYTES + W1_BYTES + NBAR * 8 + 16;
static_assert(SMEM + 2 * PATCH * PATCH * 2 + 256 <= 232448, "shared memory budget");
}  // namespace c1m

// K-major operand without swizzle: core matrices of 8 rows x 16 B; `lbo` bytes between the K chunks of one
// K = 16 step, `sbo` bytes between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version 1 (sm_100); layout type 0 = no swizzle
  return d;
}

// relu + round-to-nearest fp16 pair: low half = lo, high half = hi
__device__ __forceinline__ uint32_t pack_h2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <bool PAIR>
__global__ void __launch_bounds__(c1m::THREADS, 1)
conv1ab_mma_kernel(const __grid_constant__ CUtensorMap tmW, const Conv1abArgs p) {
  using namespace c1m;
  constexpr int WBLKP = PAIR ? WBLK / 2 : WBLK;  // bytes of one resident conv1b weight block in this CTA
  extern __shared__ uint8_t smem_raw[];
  __shared__ __half s_patch[2][PATCH * PATCH];
  __shared__ float s_bias[64];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sStage = smem;
  uint8_t *sW = sStage + 2 * STAGE;
  uint8_t *sA1 = sW + WBYTES;
  uint8_t *sW1 = sA1 + 2 * A1_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sW1 + W1_BYTES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NBAR);

  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };          // im2col producers -> MMA
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };   // conv1a MMAs done -> producers (commit)
  auto e1_done = [&](int s) { return bar0 + 8u * (4 + s); };   // conv1a epilogue -> MMA: slab s written AND conv1a accumulators drained
  auto s_empty = [&](int s) { return bar0 + 8u * (6 + s); };   // conv1b MMAs done -> conv1a epilogue (commit)
  auto t_full = [&](int s) { return bar0 + 8u * (8 + s); };    // conv1b MMAs done -> conv1b epilogue (commit)
  auto t_empty = [&](int s) { return bar0 + 8u * (10 + s); };  // conv1b epilogue -> MMA
  const uint32_t c_full = bar0 + 8u * 12;                      // conv1a MMAs done -> conv1a epilogue (commit)
  const uint32_t w_full = bar0 + 8u * 13;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++) {  // per-thread arrivals in a single CTA; one (remote) arrival per warp and rank in a pair
      mbar_init(a_full(s), PAIR ? 6 : NPROD);
      mbar_init(a_empty(s), 1);
      mbar_init(e1_done(s), PAIR ? 16 : 256);
      mbar_init(s_empty(s), 1);
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), PAIR ? 8 : 128);
    }
    mbar_init(c_full, 1);
    mbar_init(w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
  }
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  if constexpr (PAIR) cluster_sync_all();  // both CTAs' barriers exist before anything can signal across the pair
  if (warp == 2) {
    if constexpr (PAIR) { tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish(); }
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.b1b[threadIdx.x];
  if (threadIdx.x >= 64 && threadIdx.x < 64 + W1_BYTES / 16) {  // conv1a weights (hi | lo), already in operand layout
    const int i = threadIdx.x - 64;
    // a pair member keeps couts [32 * rank, +32) of each part at the start of the part (8-row groups are 256 B apart)
    const int src = PAIR ? (i / 128) * 128 + static_cast<int>(rank) * 64 + (i % 128) % 64 : i;
    if (!PAIR || (i % 128) < 64) st_shared_v4(smem_u32(sW1) + i * 16, __ldg(reinterpret_cast<const uint4 *>(p.w1m) + src));
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();  // conv2a's CTAs may take over an SM (and load their weights) as soon as this CTA leaves it
  if (warp == 0) {  // conv1b weights: resident for the CTA's lifetime (a pair member loads its 32 couts of every tap)
    if (lane == 0) {
      mbar_expect_tx(w_full, 9 * WBLKP);
      for (int wb = 0; wb < 9; wb++) tma_load_2d(smem_u32(sW + wb * WBLKP), &tmW, w_full, 0, wb * 64 + static_cast<int>(rank) * 32);
    }
    if constexpr (PAIR) mbar_wait(w_full, 0);
  }
  if constexpr (PAIR) cluster_sync_all();  // nobody starts before both halves of the weights are in

  // Work distribution (as in conv_tc.cuh): plain, CTA b takes items b, b + grid, ...; pair j = b / 2 takes item pairs
  // j, j + grid / 2, ... and rank r works on item 2q + r (clamped: an odd last item is computed by both ranks).
  const int q_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int q_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int q_end = PAIR ? (p.n_items + 1) >> 1 : p.n_items;
  const int n_mine = q_first < q_end ? (q_end - q_first + q_step - 1) / q_step : 0;
  auto item_of = [&](int k) {  // k-th item of this CTA
    const int q = q_first + k * q_step;
    return PAIR ? min(2 * q + static_cast<int>(rank), p.n_items - 1) : q;
  };
  auto decode = [&](int item, int &x0, int &y0, int &b) {
    x0 = (item % p.tiles_x) * 16;
    const int t = item / p.tiles_x;
    y0 = (t % p.tiles_y) * 16;
    b = t / p.tiles_y;
  };
  // "ready" barriers live in the leader of a pair: one remote arrival per warp.  Where shared-memory contents are
  // handed over (A1, slab), the writers run fence.proxy.async first and the arrival is the plain CTA-scope-release one,
  // exactly CUTLASS's 2-SM hand-off (ClusterBarrier::arrive after fence_view_async_shared): each rank's tensor core reads
  // its OWN shared memory, written by its own warps before the arrival left the SM.  A cluster-scope release here costs
  // 900 - 1800 cycles per arrival (traced) and puts the conv1a chain back on the critical path.
  auto arrive_ready = [&](uint32_t bar) {
    if constexpr (PAIR) {
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(bar, 0));
    } else {
      mbar_arrive(bar);
    }
  };

  if (warp == 1) {
    if (rank == 0) {
      // ---------------------------------------------- MMA issuer (warp-uniform; one elected lane issues; pair: leader only)
      constexpr uint32_t idesc = PAIR ? umma_idesc_f16_pair(64) : umma_idesc_f16(64);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
        if constexpr (PAIR) umma_f16_pair(d, a, b, idesc, acc);
        else umma_f16(d, a, b, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {
        if constexpr (PAIR) umma_commit_pair(bar);
        else umma_commit(bar);
      };
      auto wait_ready = [&](uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };
      const uint32_t sStage_u = smem_u32(sStage), sW_u = smem_u32(sW), sA1_u = smem_u32(sA1), sW1_u = smem_u32(sW1);
      auto issue_c1a = [&](uint32_t n) {  // conv1a of the n-th item: 3 M-tiles x (hi, lo); its accumulators are free (e1_done(n-1))
        const uint32_t buf = n & 1;
        wait_ready(a_full(buf), (n >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bh = umma_desc_nosw(sW1_u, 128, 256), bl = umma_desc_nosw(sW1_u + W1_PART, 128, 256);
#pragma unroll
          for (int m = 0; m < MT; m++) {
            const uint64_t a = umma_desc_nosw(sA1_u + buf * A1_BYTES + m * 4096, 128, 256);
            mma(tmem_base + C1A_COL + m * 64, a, bh, 0u);
            mma(tmem_base + C1A_COL + m * 64, a, bl, 1u);
          }
          commit(a_empty(buf));
          commit(c_full);
        }
        __syncwarp();
        C1M_T(static_cast<int>(n), 0);
      };
      mbar_wait(w_full, 0);
      if (n_mine > 0) issue_c1a(0);
      for (int n = 0; n < n_mine; n++) {
        const uint32_t st = n & 1, ph = (n >> 1) & 1;
        wait_ready(e1_done(st), ph);  // slab of item n written, conv1a accumulators free again
        C1M_T(n, 5);
        if (n + 1 < n_mine) issue_c1a(n + 1);
        wait_ready(t_empty(st), ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + st * 128;
        if (elect_one()) {
          const uint64_t a0 = umma_desc_sw128(sStage_u + st * STAGE, PW * 128);
#pragma unroll
          for (int dx = 0; dx < 3; dx++)
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
              const uint64_t b0 = umma_desc_sw128(sW_u + (dy * 3 + dx) * WBLKP, 1024);
#pragma unroll
              for (int h = 0; h < 2; h++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                  mma(d_tmem + h * 64, a0 + static_cast<uint64_t>((dy * PW + h * 8 + dx) * 8 + 2 * k), b0 + 2 * k, (dx | dy | k) ? 1u : 0u);
            }
          commit(s_empty(st));
          commit(t_full(st));
        }
        __syncwarp();
        C1M_T(n, 6);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------ conv1b epilogue: bias + ReLU + 2x2 pool -> fp16 NHWC
    const int wq = warp & 3;
    const int hl = wq * 4 + (lane >> 3), wl = lane & 7;
    for (int n = 0; n < n_mine; n++) {
      int x0, y0, b;
      decode(item_of(n), x0, y0, b);
      const uint32_t st = n & 1;
      mbar_wait(t_full(st), (n >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; h++) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + st * 128 + h * 64;
        epilogue_relu_pool<64>(taddr, s_bias, lane, hl, wl, x0 + h * 8, y0, b, 0, p.H, p.W, 64, p.out);
      }
      tc_fence_before();
      arrive_ready(t_empty(st));
    }
  } else if (warp >= 8) {
    // ------------------------------------------------ conv1a epilogue, channels 32 * half .. +32:
    // TMEM -> x 1/255 -> ReLU -> fp16 -> 16-byte chunks 4 * half .. +4 of every row of the swizzled slab
    const int wq = warp & 3, half = (warp - 8) >> 2;
    const float scale = 1.0f / 255.0f;  // cv::Mat::convertTo(CV_32FC1, 1.f / 255.f), sp_extractor.cpp:386
    const uint32_t sStage_u = smem_u32(sStage);
    for (int n = 0; n < n_mine; n++) {
      const uint32_t st = n & 1, ph = (n >> 1) & 1;
#ifdef SPFE_C1M_TRACE
      const long long pt0 = clock64();
#endif
      mbar_wait(c_full, n & 1);
      if (warp == 8) C1M_T(n, 1);
#ifdef SPFE_C1M_TRACE
      const long long pt1 = clock64();
#endif
      mbar_wait(s_empty(st), ph ^ 1);  // the conv1b MMAs that read this slab two items ago are done
      if (warp == 8) C1M_T(n, 2);
      tc_fence_after();
      const uint32_t stage_u = sStage_u + st * STAGE;
#pragma unroll 1
      for (int m = 0; m < MT; m++) {
        const int q0 = m * 128 + wq * 32;
        if (q0 >= NPIX) break;  // warp-uniform
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + C1A_COL + m * 64 + half * 32;
        float v[32];
        tmem_ld16(taddr, v);
        tmem_ld16(taddr + 16, v + 16);
        tmem_ld_wait();
        const int q = q0 + lane;
        if (q < NPIX) {
          const int hh = q / HALO, j = q - hh * HALO;
          const int row = hh * PW + j;
          const uint32_t dst = stage_u + row * 128;
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint4 o;
            o.x = pack_h2_relu(v[g * 8 + 0] * scale, v[g * 8 + 1] * scale);
            o.y = pack_h2_relu(v[g * 8 + 2] * scale, v[g * 8 + 3] * scale);
            o.z = pack_h2_relu(v[g * 8 + 4] * scale, v[g * 8 + 5] * scale);
            o.w = pack_h2_relu(v[g * 8 + 6] * scale, v[g * 8 + 7] * scale);
            st_shared_v4(dst + (((half * 4 + g) ^ (row & 7)) << 4), o);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      if (warp == 8) C1M_T(n, 3);
#ifdef SPFE_C1M_TRACE
      const long long pt2 = clock64();
#endif
      arrive_ready(e1_done(st));
      if (warp == 8) C1M_T(n, 4);
#ifdef SPFE_C1M_TRACE
      if (blockIdx.x == 1 && warp == 8 && lane == 0 && n >= 200 && n < 216) {
        g_c1m_trace[n - 200][7] = pt2 - pt1; g_c1m_trace[n - 200][8] = pt1 - pt0; g_c1m_trace[n - 200][9] = clock64() - pt2;
      }
#endif
    }
  } else {
    // ------------------------------------------------ im2col producers (warps 0, 2, 3): u8 patch -> A1 rows (9 taps, 255, zeros)
    const int ptid = (warp == 0 ? 0 : warp - 1) * 32 + lane;  // 0 .. 95
    const uint32_t sA1_u = smem_u32(sA1);
    auto row_addr = [&](int buf, int q) { return sA1_u + buf * A1_BYTES + (q >> 3) * 256 + (q & 7) * 16; };
    constexpr int PER_T = (PATCH * PATCH + NPROD - 1) / NPROD;  // patch elements per thread (400 bytes over 96 threads)
    // origin of the patch is (y0-2, x0-2); 0 outside the image == conv1a's zero padding
    auto load_patch = [&](int item, int e) -> unsigned {
      int x0, y0, b;
      decode(item, x0, y0, b);
      const int r = e / PATCH, c = e - r * PATCH;
      const int y = y0 - 2 + r, x = x0 - 2 + c;
      if (e < PATCH * PATCH && y >= 0 && y < p.H && x >= 0 && x < p.W)
        return __ldg(p.img + (static_cast<size_t>(b) * p.H + y) * p.W + x);
      return 0u;
    };
    for (int buf = 0; buf < 2; buf++)  // rows 324..383 of the third M-tile: zero once, never written again
      for (int q = NPIX + ptid; q < MT * 128; q += NPROD) {
        st_shared_v4(row_addr(buf, q), make_uint4(0u, 0u, 0u, 0u));
        st_shared_v4(row_addr(buf, q) + 128, make_uint4(0u, 0u, 0u, 0u));
      }
    unsigned raw[PER_T];
#pragma unroll
    for (int k = 0; k < PER_T; k++) raw[k] = n_mine > 0 ? load_patch(item_of(0), ptid + k * NPROD) : 0u;
    for (int n = 0; n < n_mine; n++) {
      int x0, y0, b;
      decode(item_of(n), x0, y0, b);
      const uint32_t buf = n & 1, ph = (n >> 1) & 1;
      // publish patch(n) (in flight since the previous iteration); the barrier also says everyone is done with the patch
      // that lived in this buffer two items ago
#pragma unroll
      for (int k = 0; k < PER_T; k++)
        if (ptid + k * NPROD < PATCH * PATCH) s_patch[buf][ptid + k * NPROD] = __ushort2half_rn(static_cast<unsigned short>(raw[k]));
      named_bar_sync(1, NPROD);
      if (n + 1 < n_mine) {  // next item's patch: in flight during the im2col below, published by the next iteration
        const int next = item_of(n + 1);
#pragma unroll
        for (int k = 0; k < PER_T; k++) raw[k] = load_patch(next, ptid + k * NPROD);
      }
      const unsigned short *patch = reinterpret_cast<const unsigned short *>(s_patch[buf]);
      mbar_wait(a_empty(buf), ph ^ 1);  // the conv1a MMAs that read this buffer two items ago are done
      for (int q = ptid; q < NPIX; q += NPROD) {
        const int hh = q / HALO, j = q - hh * HALO;
        const int y = y0 - 1 + hh, x = x0 - 1 + j;  // halo pixel (hh, j)
        uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = make_uint4(0u, 0u, 0u, 0u);
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
          const unsigned short *pp = patch + hh * PATCH + j;
          unsigned t[9];
#pragma unroll
          for (int k = 0; k < 9; k++) t[k] = pp[(k / 3) * PATCH + k % 3];
          c0.x = t[0] | (t[1] << 16);
          c0.y = t[2] | (t[3] << 16);
          c0.z = t[4] | (t[5] << 16);
          c0.w = t[6] | (t[7] << 16);
          c1.x = t[8] | (0x5BF8u << 16);  // k = 9: 255.0 in fp16, multiplies the bias row of W1
        }
        st_shared_v4(row_addr(buf, q), c0);
        st_shared_v4(row_addr(buf, q) + 128, c1);
      }
      fence_proxy_async_smem();
      arrive_ready(a_full(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace spfe
