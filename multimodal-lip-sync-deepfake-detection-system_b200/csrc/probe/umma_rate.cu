// Standalone micro-benchmark (not part of the library): tcgen05.mma issue/retire rate for M=128, K=16 bf16 SS-mode with the
// K-major SWIZZLE_NONE layout, as a function of N and of the alignment of the A start address (shifted tap views).
#include <cstdio>
#include <cuda_runtime.h>
#include "../umma.cuh"
using namespace umma;

__global__ void __launch_bounds__(128) rate_kernel(int N, int a_off_units, int n_mma, int distinct, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 1) {
    const uint32_t sa = smem_u32(smem) >> 4, sb = (smem_u32(smem) + 96 * 1024) >> 4;
    const uint64_t hi = ((uint64_t)8 << 32) | ((uint64_t)1 << 46);
    const uint64_t da_hi = hi | ((uint64_t)600 << 16), db_hi = hi | ((uint64_t)N << 16);
    const uint32_t idesc = idesc_bf16(128, N);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t a = sa + a_off_units + (distinct ? (uint32_t)(i & 3) * 128u : 0u);
        mma_bf16_ss(tmem_base + (uint32_t)((i & 3) * N) % 512u, da_hi | a, db_hi | sb, idesc, 1u);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one()) { out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* d; cudaMalloc(&d, 8);
  const int n_mma = 4096;
  for (int N : {32, 64, 128, 256})
    for (int off : {0, 1, 4, 25})
      for (int distinct : {0, 1}) {
        if (N * 4 > 512 && distinct) {}
        rate_kernel<<<1, 128, 180 * 1024>>>(N, off, n_mma, distinct, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d a_off=%2d units distinctA=%d : %7.1f cycles/MMA  (ideal %d)  %s\n", N, off, distinct, (double)c / n_mma, N / 2,
               cudaGetErrorString(e));
      }
  return 0;
}
