// Standalone micro-benchmark: per-SM and chip-wide throughput of 1-D cp.async.bulk global->shared copies (the producer path of
// umma_conv.cu) as a function of copy size, source alignment and number of copies in flight.
#include <cstdio>
#include <cuda_runtime.h>
#include "../umma.cuh"
using namespace umma;

__global__ void __launch_bounds__(64) bulk_kernel(const uint8_t* src, size_t src_bytes, int copy_bytes, int align_off, int dst_off, int inflight, int iters,
                                                   long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) { for (int i = 0; i < inflight; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    size_t off = ((size_t)blockIdx.x * 7919u * 4096u) & (src_bytes / 2 - 1);
    for (int it = 0; it < iters; ++it) {
      const int s = it % inflight;
      if (it >= inflight) mbar_wait(&bar[s], ((it / inflight) - 1) & 1);
      mbar_arrive_expect_tx(&bar[s], copy_bytes);
      bulk_g2s(smem + dst_off + (size_t)s * (copy_bytes + 128), src + (off & ~size_t(127)) + align_off, copy_bytes, &bar[s]);
      off = (off + 1048583u * 128u) & (src_bytes / 2 - 1);
    }
    for (int s = 0; s < inflight && s < iters; ++s) {
      const int last = ((iters - 1 - s) / inflight) * inflight + s;  // last iteration that used slot s
      (void)last;
    }
    // drain: wait for the final phase of every slot
    for (int s = 0; s < inflight; ++s) {
      int uses = (iters - s + inflight - 1) / inflight;
      if (uses > 0) mbar_wait(&bar[s], (uses - 1) & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const size_t src_bytes = 512ull << 20;
  uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  long long* d; cudaMalloc(&d, 148 * 8);
  long long h[148];
  for (int grid : {1, 148})
    for (int cb : {2048, 9024, 18432})
      for (int inflight : {4, 8})
        for (int al : {0, 16})
        for (int dof : {0, 16}) {
          if ((size_t)(cb + 128) * inflight + 128 > 190 * 1024) continue;
          const int iters = 400;
          bulk_kernel<<<grid, 64, (size_t)(cb + 128) * inflight + 128>>>(src, src_bytes, cb, al, dof, inflight, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("grid=%3d copy=%5d B inflight=%d src+%2d dst+%2d : %6.1f B/cycle/SM  (%7.0f cycles per copy)  %s\n", grid, cb, inflight, al, dof,
                 (double)cb * iters / mx, (double)mx / iters, cudaGetErrorString(e));
        }
  return 0;
}
