// Microbenchmark: sustained tcgen05.mma rate of a CTA pair (cta_group::2, M = 256 over two SMs) next to the single-CTA
// rate (tools/umma_rate.cu), operands in shared memory (128B swizzle).  Question: is the ~50-cycle floor of an
// M128 x N64 x K16 MMA the shared-memory read of A + B (6 KB at 128 B/clk)?  In a pair each SM reads its own A (4 KB) but
// only half of B, so N = 64 should drop to ~40 cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sp_orb_slam_b200/csrc -o /tmp/umma2_rate tools/umma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace spfe;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

template <int N, int UNROLL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) rate2_kernel(long long *out, int n_mma, int distinct_a) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += 256) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  cluster_sync_all();
  const uint32_t tm = tslot;
  if (warp == 1) {
    if (rank == 0) {
      const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
      constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += UNROLL) {
        const uint64_t a = umma_desc_sw128(a_base + ((i / UNROLL) % distinct_a) * 16384);
        const uint64_t b = umma_desc_sw128(b_base + ((i / UNROLL) % 2) * (N / 2 * 128));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < UNROLL; k++) umma2_f16(tm + ((i / UNROLL) & 1) * 256, a + 2 * (k & 3) + 64 * (k >> 2), b + 2 * (k & 3), idesc, 1);
        }
        __syncwarp();
      }
      if (elect_one()) umma2_commit_mc(smem_u32(&bar), 3);
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0);
      const long long t1 = clock64();
      if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    } else {
      mbar_wait(smem_u32(&bar), 0);
    }
  }
  tc_fence_before(); __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
  }
}

template <int N, int UNROLL>
void run(int distinct_a) {
  long long *d; cudaMalloc(&d, 8);
  const int smem = 1024 + 200 * 1024;
  cudaFuncSetAttribute(rate2_kernel<N, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 8192;
  for (int rep = 0; rep < 2; rep++) rate2_kernel<N, UNROLL><<<148, 256, smem>>>(d, n, distinct_a);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("cta_group::2 M=256 N=%3d unroll=%2d : %.1f cycles per MMA (= per 128 rows on each SM; single-CTA floor max(~50, N/2))  %s\n",
         N, UNROLL, double(h) / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 16>(6); run<64, 4>(6); run<128, 16>(6); run<256, 16>(6); run<32, 16>(6);
  return 0;
}
