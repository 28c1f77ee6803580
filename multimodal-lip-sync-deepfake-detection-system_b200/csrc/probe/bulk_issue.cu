// Standalone micro-benchmark: does the ~490-cycle cost of one cp.async.bulk issue serialise per thread, per warp or per SM?
// W warps x L lanes issue copies concurrently (one mbarrier slot ring per issuing thread).
#include <cstdio>
#include <cuda_runtime.h>
#include "../umma.cuh"
using namespace umma;

__global__ void __launch_bounds__(256) issue_kernel(const uint8_t* src, int copy_bytes, int warps, int lanes, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[64][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) for (int j = 0; j < 4; ++j) mbar_init(&bar[i][j], 1); fence_barrier_init(); }
  __syncthreads();
  if (warp < warps && lane < lanes) {
    const int id = warp * lanes + lane;
    uint8_t* dst = smem + (size_t)id * 4 * copy_bytes;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it & 3;
      if (it >= 4) mbar_wait(&bar[id][s], ((it >> 2) - 1) & 1);
      mbar_arrive_expect_tx(&bar[id][s], copy_bytes);
      bulk_g2s(dst + (size_t)s * copy_bytes, src + ((size_t)(id * 131 + it * 17) % 4096) * 4096, copy_bytes, &bar[id][s]);
    }
    for (int s = 0; s < 4; ++s) { int uses = (iters - s + 3) / 4; if (uses > 0) mbar_wait(&bar[id][s], (uses - 1) & 1); }
    cycles[id] = clock64() - t0;
  }
}

int main() {
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  uint8_t* src; cudaMalloc(&src, 64ull << 20); cudaMemset(src, 1, 64ull << 20);
  long long* d; cudaMalloc(&d, 64 * 8);
  long long h[64];
  const int cb = 2048, iters = 400;
  for (int warps : {1, 2, 4, 8})
    for (int lanes : {1, 2, 4}) {
      if (warps * lanes * 4 * cb > 190 * 1024) continue;
      issue_kernel<<<1, 256, (size_t)warps * lanes * 4 * cb>>>(src, cb, warps, lanes, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, warps * lanes * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < warps * lanes; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("warps=%d lanes=%d : %7.1f cycles per copy per SM (%d issuers, %d B copies)  %s\n", warps, lanes,
             (double)mx / (iters * warps * lanes), warps * lanes, cb, cudaGetErrorString(e));
    }
  return 0;
}
