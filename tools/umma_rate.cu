// Microbenchmark: sustained tcgen05.mma rate (cycles per MMA) for M=128, K=16, N in {64,128,256},
// operands in shared memory (128B swizzle layout), one CTA per SM, optional concurrent st.shared traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sp_orb_slam_b200/csrc -o /tmp/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace spfe;

template <int N, int UNROLL, int M>
__global__ void __launch_bounds__(256, 1) rate_kernel(long long *out, int n_mma, int distinct_a, int writers) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += 256) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 2) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tslot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
    constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
    t0 = clock64();
    for (int i = 0; i < n_mma; i += UNROLL) {
      const uint64_t a = umma_desc_sw128(a_base + ((i / UNROLL) % distinct_a) * 16384);
      const uint64_t b = umma_desc_sw128(b_base + ((i / UNROLL) % 2) * (N * 128));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < UNROLL; k++) umma_f16(tm + ((i / UNROLL) & 1) * 256, a + 2 * (k & 3) + 64 * (k >> 2), b + 2 * (k & 3), idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 4 && writers) {
    // concurrent shared-memory store traffic (like TMA / producer writes) into a scratch region
    uint4 v = make_uint4(1, 2, 3, 4);
    const uint32_t dst = smem_u32(smem + 128 * 1024) + (threadIdx.x - 128) * 16;
    for (int i = 0; i < writers; i++) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + (i & 7) * 2048), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int UNROLL, int M>
void run(int distinct_a, int writers) {
  long long *d; cudaMalloc(&d, 8);
  const int smem = 1024 + 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<N, UNROLL, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 8192;
  for (int rep = 0; rep < 2; rep++) rate_kernel<N, UNROLL, M><<<148, 256, smem>>>(d, n, distinct_a, writers);
  cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("M=%3d N=%3d unroll=%2d distinctA=%d writers=%6d : %.1f cycles/MMA (ideal %d)  %s\n", M, N, UNROLL, distinct_a, writers, double(h) / n, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 4, 128>(6, 0); run<64, 12, 128>(6, 0); run<64, 16, 128>(6, 0);
  run<128, 4, 128>(6, 0); run<128, 16, 128>(6, 0);
  run<256, 4, 128>(6, 0); run<256, 16, 128>(6, 0);
  run<64, 16, 64>(6, 0); run<128, 16, 64>(6, 0); run<256, 16, 64>(6, 0);
  run<32, 16, 128>(6, 0); run<16, 16, 128>(6, 0);
  return 0;
}
