// Correctness probe: can a K-major SWIZZLE_128B UMMA A-descriptor start at a 128-byte row that is NOT 1024-byte
// aligned (start = slab + dx*128, SBO = row pitch), and what must base_offset be?  D = A * I, so D shows exactly
// which shared-memory rows / chunks the tensor core read.
#include <cstdio>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace spfe;

constexpr int PITCH = 16;            // slab pixels per image row (multiple of 8)
constexpr int ROWS = 18;             // image rows in the slab
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(float *out, int dx, int dy, int base_off) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __half *sA = reinterpret_cast<__half *>(smem);                       // slab: ROWS*PITCH pixels x 64 ch, SW128 by absolute address
  __half *sB = reinterpret_cast<__half *>(smem + ROWS * PITCH * 128);  // 64 x 64 identity, SW128
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < ROWS * PITCH * 64; i += 128) {
    const int p = i / 64, c = i % 64;                                  // pixel p, channel c: value = (p * 7 + c) % 251 (exact in fp16)
    const int chunk = c / 8, phys = p * 128 + (((chunk ^ (p & 7)) << 4)) + (c % 8) * 2;
    *reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(sA) + phys) = __float2half(static_cast<float>((p * 7 + c) % 251));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const int chunk = k / 8, phys = n * 128 + (((chunk ^ (n & 7)) << 4)) + (k % 8) * 2;
    *reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(sB) + phys) = __float2half(n == k ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 64); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tslot;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint32_t a_addr = smem_u32(sA) + (dy * PITCH + dx) * 128;
      const uint32_t idesc = umma_idesc_f16(64);
      for (int k = 0; k < 4; k++)
        umma_f16(tm, desc(a_addr, PITCH * 128, base_off) + 2 * k, desc(smem_u32(sB), 1024, 0) + 2 * k, idesc, k ? 1u : 0u);
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tmem_ld16(tm + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

int main() {
  float *d; cudaMalloc(&d, 128 * 64 * 4);
  std::vector<float> h(128 * 64);
  const int smem = 1024 + ROWS * PITCH * 128 + 64 * 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int dy = 0; dy < 3; dy += 2)
    for (int dx = 0; dx < 4; dx++)
      for (int bo : {0, dx, (8 - dx) & 7}) {
        cudaMemset(d, 0, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(d, dx, dy, bo);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; m++)
          for (int n = 0; n < 64; n++) {
            const int p = (m / 8 + dy) * PITCH + (m % 8) + dx;
            if (h[m * 64 + n] != static_cast<float>((p * 7 + n) % 251)) bad++;
          }
        printf("dy=%d dx=%d base_offset=%d : %d / 8192 wrong %s\n", dy, dx, bo, bad, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
