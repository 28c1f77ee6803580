// Device helpers shared by the fused token-path kernels (tok_fused.cu, tok_front.cu): fp16 operand planes in shared memory
// (K-major core-matrix layout: planes of 8 channels, rows 16 B apart, no swizzle), TMEM stores, exact-erf GELU.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "umma.cuh"

namespace lsd {
namespace tokc {

using namespace umma;

constexpr uint32_t PLANE = 2048;                     // one 8-channel plane of a 128-row operand

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // D = f32, A = B = f16, K-major, dense
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
// 8 consecutive fp32 -> one 16-byte fp16 row piece of a plane
__device__ __forceinline__ void st_plane8(uint32_t addr, const float* v) {
  st_shared_v4(addr, pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
}
// v[0..32) += p[0..32)  (p 16-byte aligned, read-only parameters: vector loads through the read-only path)
__device__ __forceinline__ void add32(float* v, const float* p) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + e);
    v[4 * e] += t.x; v[4 * e + 1] += t.y; v[4 * e + 2] += t.z; v[4 * e + 3] += t.w;
  }
}
// Exact-erf GELU (approximate='none', temporal.py:39-49 / nn.TransformerEncoderLayer activation="gelu") with erf from
// Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): two MUFU (rcp, ex2) + 8 FMA instead of erff's ~30 instructions — the FFN
// epilogue (1024 activations per token and layer) was the largest compute phase of the kernel.  The result is rounded to fp16.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float w = poly * t * __expf(-z * z);          // = 1 - erf(|x| / sqrt 2) = erfc
  return 0.5f * x * (x >= 0.f ? 2.0f - w : w);         // 1 + erf(x / sqrt 2), without cancellation for x < 0
}


}  // namespace tokc
}  // namespace lsd
