// Fused conv1a + conv1b + ReLU + 2x2 max-pool (reference sp_extractor.cpp:81-83, fed by :386-390).
//
// conv1b is 43 % of the network's FLOPs and its input -- conv1a's 64-channel
// full-resolution activation -- is 46 MB per 752x480 frame.  Instead of
// materialising it, every persistent conv1b CTA recomputes conv1a for the halo
// of its 8x16-pixel tile on the CUDA cores (fp32 FFMA: conv1a is the
// precision-critical layer and K = 9 is not a tensor-core shape) and writes
// the result as fp16 straight into the 128B-swizzled shared-memory slabs that
// the conv1b tcgen05 MMAs read.  The u8 image is the only tensor read from HBM.
//
// Warps: 0 = conv1b weight TMA (once), 1 = MMA issuer, 2 = TMEM allocator,
// 4-7 = epilogue, 8-15 = conv1a producers.  A stage = the three dx-shifted
// slabs of one tile (3 x 18 rows x 8 px x 64 ch); two stages and two TMEM
// accumulators, so producers, tensor core and epilogue work on three different
// tiles at once.
#pragma once
#include "conv_tc.cuh"

namespace spfe {

struct Conv1abArgs {
  const uint8_t *img;  // [B][H][W]
  const float *w1a;    // [9][64]
  const float *b1a;    // [64]
  const float *b1b;    // [64]
  __half *out;         // [B][H/2][W/2][64]
  int B, H, W, tiles_x, tiles_y, n_items;
};

namespace c1ab {
constexpr int SLAB = 18 * 1024, STAGE = 3 * SLAB, WBLK = 64 * 128, WBYTES = 9 * WBLK;
constexpr int PATCH_H = 20, PATCH_W = 12, HALO_H = 18, HALO_W = 10;
constexpr int THREADS = 512, PRODUCERS = 256;
constexpr int SMEM = 1024 + 2 * STAGE + WBYTES + 9 * 8 + 16;  // dynamic part (patch + bias are static)
}  // namespace c1ab

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4 &v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(c1ab::THREADS, 1)
conv1ab_kernel(const __grid_constant__ CUtensorMap tmW, const Conv1abArgs p) {
  using namespace c1ab;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_patch[2][PATCH_H * PATCH_W];  // static: keeps the accesses LDS (not generic LD)
  __shared__ float s_bias[64];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sStage = smem;
  uint8_t *sW = sStage + 2 * STAGE;
  float *sBias = s_bias;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sW + WBYTES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);

  const uint32_t bar0 = smem_u32(bars);
  auto s_full = [&](int s) { return bar0 + 8u * s; };         // producers -> MMA   (256 arrivals)
  auto s_empty = [&](int s) { return bar0 + 8u * (2 + s); };  // MMA -> producers   (tcgen05.commit)
  auto t_full = [&](int s) { return bar0 + 8u * (4 + s); };   // MMA -> epilogue
  auto t_empty = [&](int s) { return bar0 + 8u * (6 + s); };  // epilogue -> MMA    (128 arrivals)
  const uint32_t w_full = bar0 + 8u * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(s_full(s), PRODUCERS);
      mbar_init(s_empty(s), 1);
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 128);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmW);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), 128);
    tmem_relinquish();
  }
  if (threadIdx.x < 64) sBias[threadIdx.x] = p.b1b[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int &x0, int &y0, int &b) {
    x0 = (item % p.tiles_x) * 8;
    const int t = item / p.tiles_x;
    y0 = (t % p.tiles_y) * 16;
    b = t / p.tiles_y;
  };

  if (warp == 0) {
    if (lane == 0) {  // conv1b weights: resident for the CTA's lifetime
      mbar_expect_tx(w_full, WBYTES);
      for (int wb = 0; wb < 9; wb++) tma_load_2d(smem_u32(sW + wb * WBLK), &tmW, w_full, 0, wb * 64);
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform, one elected lane issues)
    constexpr uint32_t idesc = umma_idesc_f16(64);
    mbar_wait(w_full, 0);
    const uint32_t sStage_u = smem_u32(sStage), sW_u = smem_u32(sW);
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, tcount++) {
      const uint32_t st = tcount & 1, ph = (tcount >> 1) & 1;
      mbar_wait(t_empty(st), ph ^ 1);
      mbar_wait(s_full(st), ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + st * 64;
      if (elect_one()) {
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
          const uint64_t a0 = umma_desc_sw128(sStage_u + st * STAGE + dx * SLAB);
#pragma unroll
          for (int dy = 0; dy < 3; dy++) {
            const uint64_t a_desc = a0 + static_cast<uint64_t>(dy * (1024 >> 4));
            const uint64_t b_desc = umma_desc_sw128(sW_u + (dy * 3 + dx) * WBLK);
#pragma unroll
            for (int k = 0; k < 4; k++) umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (dx | dy | k) ? 1u : 0u);
          }
        }
        umma_commit(s_empty(st));
        umma_commit(t_full(st));
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------ epilogue: bias + ReLU + 2x2 pool -> fp16 NHWC
    const int wq = warp & 3;
    const int hl = wq * 4 + (lane >> 3), wl = lane & 7;
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, tcount++) {
      int x0, y0, b;
      decode(item, x0, y0, b);
      const uint32_t st = tcount & 1;
      mbar_wait(t_full(st), (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + st * 64;
      epilogue_relu_pool<64>(taddr, sBias, lane, hl, wl, x0, y0, b, 0, p.H, p.W, 64, p.out);
      tc_fence_before();
      mbar_arrive(t_empty(st));
    }
  } else if (warp >= 8) {
    // ------------------------------------------------ conv1a producers (fp32 FFMA -> fp16 swizzled slabs)
    // Thread = (pixel slot 0..31, 8-channel group g).  The halo has 18 x 10 pixels = 180 slots x 8 groups
    // = 1440 work items, 6 rounds of 256.  The group's 9 x 8 weights stay in registers; the image patch of
    // the NEXT tile is prefetched into a register while this tile is computed (double-buffered patch).
    const int ptid = threadIdx.x - 256;
    const int g = ptid & 7;  // 8 output channels == 16-byte chunk g of every 128-byte pixel row
    float w[9][8], bs[8];
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const float4 lo = *reinterpret_cast<const float4 *>(p.w1a + t * 64 + g * 8);
      const float4 hi = *reinterpret_cast<const float4 *>(p.w1a + t * 64 + g * 8 + 4);
      w[t][0] = lo.x; w[t][1] = lo.y; w[t][2] = lo.z; w[t][3] = lo.w;
      w[t][4] = hi.x; w[t][5] = hi.y; w[t][6] = hi.z; w[t][7] = hi.w;
    }
#pragma unroll
    for (int c = 0; c < 8; c++) bs[c] = p.b1a[g * 8 + c];
    const float scale = 1.0f / 255.0f;  // cv::Mat::convertTo(CV_32FC1, 1.f / 255.f)
    const uint32_t sStage_u = smem_u32(sStage);
    const int pr = ptid / PATCH_W, pc = ptid - pr * PATCH_W;  // this thread's patch element (ptid < 240)
    const int hh0 = (ptid >> 3) / HALO_W, j0 = (ptid >> 3) - hh0 * HALO_W;
    auto load_patch = [&](int item) -> unsigned {  // raw byte; patch origin (y0-2, x0-2); 0 outside == zero padding
      int x0, y0, b;
      decode(item, x0, y0, b);
      const int y = y0 - 2 + pr, x = x0 - 2 + pc;
      if (ptid < PATCH_H * PATCH_W && y >= 0 && y < p.H && x >= 0 && x < p.W)
        return __ldg(p.img + (static_cast<size_t>(b) * p.H + y) * p.W + x);
      return 0u;
    };
    if (blockIdx.x < p.n_items && ptid < PATCH_H * PATCH_W) s_patch[0][ptid] = static_cast<float>(load_patch(blockIdx.x)) * scale;
    named_bar_sync(1, PRODUCERS);
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, tcount++) {
      int x0, y0, b;
      decode(item, x0, y0, b);
      const uint32_t st = tcount & 1, ph = (tcount >> 1) & 1;
      const int next = item + gridDim.x;
      const unsigned raw_next = next < p.n_items ? load_patch(next) : 0u;  // in flight during the compute below
      const float *patch = s_patch[st];
      mbar_wait(s_empty(st), ph ^ 1);  // the MMAs of the tile that used this stage two tiles ago are done
      const uint32_t stage_u = sStage_u + st * STAGE;
#pragma unroll 2
      for (int k = 0; k < 6; k++) {
        if (k == 5 && ptid >= HALO_H * HALO_W * 8 - 5 * PRODUCERS) break;
        int j = j0 + 2 * k, hh = hh0 + 3 * k;  // pixel slot + 32k  ->  (+3 rows, +2 columns) with one carry
        if (j >= HALO_W) { j -= HALO_W; hh++; }
        const int y = y0 - 1 + hh, x = x0 - 1 + j;  // halo pixel
        uint4 o = make_uint4(0u, 0u, 0u, 0u);        // outside the image: conv1b's zero padding
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
          float acc[8];
#pragma unroll
          for (int c = 0; c < 8; c++) acc[c] = bs[c];
          const float *pp = patch + hh * PATCH_W + j;
#pragma unroll
          for (int t = 0; t < 9; t++) {
            const float xin = pp[(t / 3) * PATCH_W + t % 3];
#pragma unroll
            for (int c = 0; c < 8; c++) acc[c] = fmaf(w[t][c], xin, acc[c]);
          }
          o.x = pack_h2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f));
          o.y = pack_h2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
          o.z = pack_h2(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f));
          o.w = pack_h2(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
        }
        // halo column j is column (j - dx) of slab dx; row = hh*8 + col, 16-byte chunk g XOR-swizzled by row & 7
        const uint32_t base = stage_u + hh * 1024;
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
          const int col = j - dx;
          if (col >= 0 && col < 8) st_shared_v4(base + dx * SLAB + col * 128 + ((g ^ col) << 4), o);
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(s_full(st));
      if (ptid < PATCH_H * PATCH_W) s_patch[st ^ 1][ptid] = static_cast<float>(raw_next) * scale;
      named_bar_sync(1, PRODUCERS);  // next patch visible; everyone is done with this one
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace spfe
