// Standalone probe (not part of the library): pins the semantics of the K-major SWIZZLE_NONE shared-memory
// descriptor (LBO / SBO / unaligned start / overlapping "Toeplitz" views) that umma_conv.cu relies on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../umma.cuh"

using namespace umma;

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_img, uint32_t a_bytes, const uint8_t* b_img, uint32_t b_bytes,
                                                    uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep,
                                                    uint32_t b_off, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                                                    int N, int ksteps, uint32_t tmem_cols, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) & ~1023u);
  for (uint32_t i = threadIdx.x * 16; i < a_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (uint32_t i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(b_img + i);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tmem_base, tmem_cols); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_bf16(128, N);
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t da = smem_desc(smem_u32(sa) + a_off + k * a_kstep, a_lbo, a_sbo);
      const uint64_t db = smem_desc(smem_u32(sb) + b_off + k * b_kstep, b_lbo, b_sbo);
      mma_bf16_ss(tb, da, db, idesc, k > 0 ? 1u : 0u);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, tmem_cols);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

struct Case { const char* name; int N, K; int variant; bool swap; };

int main() {
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<Case> cases = {
      {"canonical N=64 K=64", 64, 64, 1, false},  // (the LBO/SBO-swapped variant faults: measured once, removed)
      {"canonical N=256 K=128", 256, 128, 1, false}, {"halo-view N=64 K=64 (unaligned start, SBO=160)", 64, 64, 2, false},
      {"halo-view N=128 K=64", 128, 64, 2, false},   {"toeplitz N=64 K=32 (LBO=16, overlapping rows)", 64, 32, 3, false},
      {"halo-view stride-2 groups N=64 K=64 (SBO=320)", 64, 64, 4, false},
  };
  int fails = 0;
  for (const Case& c : cases) {
    const int M = 128, N = c.N, K = c.K;
    srand(1234 + N + K + c.variant);
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    std::vector<uint8_t> aimg, bimg;
    uint32_t a_off = 0, a_lbo = 0, a_sbo = 0, a_kstep = 0;
    auto rnd = []() { return bf2f(f2bf((float)(rand() % 2001 - 1000) / 1000.0f)); };
    for (auto& x : B) x = rnd();
    // B: canonical [k-chunk][n][8 elem]
    bimg.assign((size_t)N * K * 2, 0);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const uint16_t h = f2bf(B[(size_t)n * K + k]);
        memcpy(&bimg[((size_t)(k / 8) * N + n) * 16 + (k % 8) * 2], &h, 2);
      }
    const uint32_t b_lbo = N * 16, b_sbo = 128, b_kstep = 2 * N * 16;
    if (c.variant == 1) {
      for (auto& x : A) x = rnd();
      aimg.assign((size_t)M * K * 2, 0);
      for (int r = 0; r < M; ++r)
        for (int k = 0; k < K; ++k) {
          const uint16_t h = f2bf(A[(size_t)r * K + k]);
          memcpy(&aimg[((size_t)(k / 8) * M + r) * 16 + (k % 8) * 2], &h, 2);
        }
      a_lbo = M * 16; a_sbo = 128; a_kstep = 2 * M * 16;
    } else if (c.variant == 2 || c.variant == 4) {
      // halo: P positions x K channels stored [chunk][pos][8]; row (g, r) <-> position base + g*pitch + r
      const int P = 400, base = 3, pitch = (c.variant == 2) ? 10 : 20;
      std::vector<float> X((size_t)P * K);
      for (auto& x : X) x = rnd();
      aimg.assign((size_t)P * K * 2, 0);
      for (int p = 0; p < P; ++p)
        for (int k = 0; k < K; ++k) {
          const uint16_t h = f2bf(X[(size_t)p * K + k]);
          memcpy(&aimg[((size_t)(k / 8) * P + p) * 16 + (k % 8) * 2], &h, 2);
        }
      for (int r = 0; r < M; ++r)
        for (int k = 0; k < K; ++k) A[(size_t)r * K + k] = X[(size_t)(base + (r / 8) * pitch + (r % 8)) * K + k];
      a_off = base * 16; a_lbo = P * 16; a_sbo = pitch * 16; a_kstep = 2 * P * 16;
    } else {
      // toeplitz: linear bf16 signal x; A(r, k) = x[8 r + k]
      const int L = 8 * M + K + 64;
      std::vector<float> X(L);
      for (auto& x : X) x = rnd();
      aimg.assign((size_t)((L * 2 + 15) / 16) * 16, 0);
      for (int i = 0; i < L; ++i) { const uint16_t h = f2bf(X[i]); memcpy(&aimg[(size_t)i * 2], &h, 2); }
      for (int r = 0; r < M; ++r)
        for (int k = 0; k < K; ++k) A[(size_t)r * K + k] = X[8 * r + k];
      a_off = 0; a_lbo = 16; a_sbo = 128; a_kstep = 32;
    }
    if (c.swap) std::swap(a_lbo, a_sbo);
    std::vector<float> ref((size_t)M * N, 0.f);
    for (int r = 0; r < M; ++r)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[(size_t)r * K + k] * B[(size_t)n * K + k];
        ref[(size_t)r * N + n] = (float)s;
      }
    uint8_t *da, *db; float* dD;
    cudaMalloc(&da, aimg.size()); cudaMalloc(&db, bimg.size()); cudaMalloc(&dD, ref.size() * 4);
    cudaMemcpy(da, aimg.data(), aimg.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bimg.data(), bimg.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, ref.size() * 4);
    uint32_t cols = 32; while ((int)cols < N) cols *= 2;
    const size_t smem = ((aimg.size() + 1023) & ~size_t(1023)) + bimg.size() + 1024;
    probe_kernel<<<1, 128, smem>>>(da, (uint32_t)aimg.size(), db, (uint32_t)bimg.size(), a_off, a_lbo, a_sbo, a_kstep, 0, b_lbo, b_sbo,
                                   b_kstep, N, K / 16, cols, dD);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(ref.size());
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0; for (size_t i = 0; i < ref.size(); ++i) mx = fmax(mx, fabs((double)out[i] - ref[i]));
    const bool expect_ok = !c.swap;
    const bool ok = (e == cudaSuccess) && (mx < 1e-2);
    printf("%-55s smem=%zu err=%s max|d|=%.4g -> %s\n", c.name, smem, cudaGetErrorString(e), mx, ok ? "MATCH" : "MISMATCH");
    if (ok != expect_ok) ++fails;
    cudaFree(da); cudaFree(db); cudaFree(dD);
    if (e != cudaSuccess) { printf("CUDA error, stopping\n"); return 2; }
  }
  printf(fails ? "PROBE: %d unexpected results\n" : "PROBE: all as expected\n", fails);
  return fails ? 1 : 0;
}
