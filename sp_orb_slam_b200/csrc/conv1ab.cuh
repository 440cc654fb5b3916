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
// 4-7 = epilogue, 8-15 = conv1a producers.  A stage = the halo slab of one pair
// of tiles (18 rows x 24 px x 64 ch, each halo pixel stored once); two stages
// and two TMEM accumulator sets, so producers, tensor core and epilogue work on
// three different items at once.
#pragma once
#include "conv_tc.cuh"

namespace spfe {

struct Conv1abArgs {
  const uint8_t *img;  // [B][H][W]
  const float *w1a;    // [9][64]
  const float *b1a;    // [64]
  const void *w1m;     // conv1ab_mma.cuh: conv1a weights + bias as hi | lo fp16 UMMA operands (4096 B)
  const float *b1b;    // [64]
  __half *out;         // [B][H/2][W/2][64]
  int B, H, W, tiles_x, tiles_y, n_items;
};

namespace c1ab {
// A work item is a pair of 8x16-pixel tiles (16 x 16 outputs).  Its conv1a halo is 18 x 18 pixels, stored once in a
// slab of 18 rows x 24 pixels (pitch padded to a multiple of 8 pixels) x 128 B; tap (dy, dx) of half h reads it at
// +((dy*24 + h*8 + dx) * 128) bytes (the 128B swizzle acts on absolute address bits, see conv_tc.cuh).
constexpr int PW = 24, HALO = 18, STAGE = HALO * PW * 128, WBLK = 64 * 128, WBYTES = 9 * WBLK;
constexpr int PATCH = HALO + 2;  // image patch side (conv1a's own 3x3 support)
constexpr int THREADS = 512, PRODUCERS = 256;
constexpr int ITEMS = HALO * HALO * 8;  // (halo pixel, 8-channel group) work items per tile pair
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
  __shared__ float s_patch[2][PATCH * PATCH];  // static: keeps the accesses LDS (not generic LD)
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
    tmem_alloc(smem_u32(tmem_slot), 256);  // 2 stages x 2 halves x 64 columns
    tmem_relinquish();
  }
  if (threadIdx.x < 64) sBias[threadIdx.x] = p.b1b[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int &x0, int &y0, int &b) {
    x0 = (item % p.tiles_x) * 16;
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
    // ------------------------------------------------ MMA issuer (warp-uniform, one elected lane issues 72 MMAs)
    constexpr uint32_t idesc = umma_idesc_f16(64);
    mbar_wait(w_full, 0);
    const uint32_t sStage_u = smem_u32(sStage), sW_u = smem_u32(sW);
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, tcount++) {
      const uint32_t st = tcount & 1, ph = (tcount >> 1) & 1;
      mbar_wait(t_empty(st), ph ^ 1);
      mbar_wait(s_full(st), ph);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + st * 128;
      if (elect_one()) {
        const uint64_t a0 = umma_desc_sw128(sStage_u + st * STAGE, PW * 128);
#pragma unroll
        for (int dx = 0; dx < 3; dx++)
#pragma unroll
          for (int dy = 0; dy < 3; dy++) {
            const uint64_t b0 = umma_desc_sw128(sW_u + (dy * 3 + dx) * WBLK, 1024);
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
              for (int k = 0; k < 4; k++)
                umma_f16(d_tmem + h * 64, a0 + static_cast<uint64_t>((dy * PW + h * 8 + dx) * 8 + 2 * k), b0 + 2 * k, idesc,
                         (dx | dy | k) ? 1u : 0u);
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
#pragma unroll 1
      for (int h = 0; h < 2; h++) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + st * 128 + h * 64;
        epilogue_relu_pool<64>(taddr, sBias, lane, hl, wl, x0 + h * 8, y0, b, 0, p.H, p.W, 64, p.out);
      }
      tc_fence_before();
      mbar_arrive(t_empty(st));
    }
  } else if (warp >= 8) {
    // ------------------------------------------------ conv1a producers (fp32 FFMA -> fp16 swizzled slab)
    // Thread = (pixel slot 0..31, 8-channel group g).  The halo has 18 x 18 pixels x 8 groups = 2592 work items,
    // 10 rounds of 256 plus 32.  The group's 9 x 8 weights stay in registers; the image patch of the NEXT item is
    // prefetched into registers while this item is computed (double-buffered patch, one named barrier per item).
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
    const int hh0 = (ptid >> 3) / HALO, j0 = (ptid >> 3) - hh0 * HALO;
    // patch element(s) of this thread: 400 bytes over 256 threads (origin (y0-2, x0-2); 0 outside == zero padding)
    auto load_patch = [&](int item, int e) -> unsigned {
      int x0, y0, b;
      decode(item, x0, y0, b);
      const int r = e / PATCH, c = e - r * PATCH;
      const int y = y0 - 2 + r, x = x0 - 2 + c;
      if (e < PATCH * PATCH && y >= 0 && y < p.H && x >= 0 && x < p.W)
        return __ldg(p.img + (static_cast<size_t>(b) * p.H + y) * p.W + x);
      return 0u;
    };
    if (blockIdx.x < p.n_items) {
      s_patch[0][ptid] = static_cast<float>(load_patch(blockIdx.x, ptid)) * scale;
      if (ptid + 256 < PATCH * PATCH) s_patch[0][ptid + 256] = static_cast<float>(load_patch(blockIdx.x, ptid + 256)) * scale;
    }
    named_bar_sync(1, PRODUCERS);
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, tcount++) {
      int x0, y0, b;
      decode(item, x0, y0, b);
      const uint32_t st = tcount & 1, ph = (tcount >> 1) & 1;
      const int next = item + gridDim.x;
      const bool has_next = next < p.n_items;
      const unsigned raw0 = has_next ? load_patch(next, ptid) : 0u;  // in flight during the compute below
      const unsigned raw1 = has_next ? load_patch(next, ptid + 256) : 0u;
      const float *patch = s_patch[st];
      mbar_wait(s_empty(st), ph ^ 1);  // the MMAs of the item that used this stage two items ago are done
      const uint32_t stage_u = sStage_u + st * STAGE;
      int j = j0, hh = hh0;
#pragma unroll 2
      for (int k = 0; k < 11; k++) {
        if (k == 10 && ptid >= ITEMS - 10 * PRODUCERS) break;
        const int y = y0 - 1 + hh, x = x0 - 1 + j;  // halo pixel (hh, j)
        uint4 o = make_uint4(0u, 0u, 0u, 0u);        // outside the image: conv1b's zero padding
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
          float acc[8];
#pragma unroll
          for (int c = 0; c < 8; c++) acc[c] = bs[c];
          const float *pp = patch + hh * PATCH + j;
#pragma unroll
          for (int t = 0; t < 9; t++) {
            const float xin = pp[(t / 3) * PATCH + t % 3];
#pragma unroll
            for (int c = 0; c < 8; c++) acc[c] = fmaf(w[t][c], xin, acc[c]);
          }
          o.x = pack_h2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f));
          o.y = pack_h2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
          o.z = pack_h2(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f));
          o.w = pack_h2(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
        }
        const int row = hh * PW + j;  // 16-byte chunk g of that row, XOR-swizzled by (row & 7)
        st_shared_v4(stage_u + row * 128 + ((g ^ (row & 7)) << 4), o);
        j += 32 - HALO;  // next pixel slot: +32 pixels == +1 row +14 columns (HALO = 18)
        hh += 1;
        if (j >= HALO) { j -= HALO; hh++; }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(s_full(st));
      s_patch[st ^ 1][ptid] = static_cast<float>(raw0) * scale;
      if (ptid + 256 < PATCH * PATCH) s_patch[st ^ 1][ptid + 256] = static_cast<float>(raw1) * scale;
      named_bar_sync(1, PRODUCERS);  // next patch visible; everyone is done with this one
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace spfe
