// Temporal-ring 3x3x3 convolution (64 -> 64 channels) for sm_100a: see conv_ring.cuh and stem_ring.cu for the scheme.
//
// An accumulator slot walks one 128-position column chunk along t with a ring of four 64-column blocks in TMEM; at step t every
// (K16 chunk, spatial tap) is ONE MMA with B = [W(dt=-1) | W(dt=0) | W(dt=+1)] (N = 192): the A operand — the shifted view of the
// chunk's region in slab t — is read once for the three temporal taps, which makes the 64-channel stage tensor-bound (96 cycles of
// tensor work per 4 + 6 KB of operand reads) instead of bound by the shared-memory read port (3 x 48 cycles for three N = 64 MMAs).
// Differences from the ring stem: activations are planar (two 8-channel planes per K16 chunk: LBO = region length), and the weights
// (221 KB for N = 192) do not fit next to the accumulator regions — they stream through a two-stage ring, one K16 chunk (54 KB, nine
// taps) per stage, shared by the two slots of the CTA.
//
// Warp roles (384 threads): warp 0 activation producer, warps 1-2 MMA issuers (one per slot), warp 3 TMEM allocation + weight
// producer, warps 4-11 epilogue (4-7 slot 0, 8-11 slot 1; warp % 4 = TMEM lane quarter): + bias (by MMA) (+ residual) -> ReLU ->
// bf16 -> plain or parity-split planar store.
#include "conv_ring.cuh"

#include <cstdio>

#include "lsd_kernels.h"
#include "umma.cuh"

namespace lsd {

using namespace umma;

namespace {

constexpr int CR_THREADS = 384, CR_NST = 4;

__device__ __forceinline__ uint32_t cr_div(uint32_t n, uint32_t m, int s) { return (uint32_t)(((uint64_t)n * m) >> (31 + s)); }
__device__ __forceinline__ bool cr_decode(const UcGeom& g, int64_t P, int& n, int& t, int& h, int& w) {
  if (P < 0 || P >= g.P_total) return false;
  const uint32_t Pu = (uint32_t)P;
  const uint32_t S = cr_div(Pu, g.mSL, g.sSL);
  const int r = (int)(Pu - S * (uint32_t)g.SL);
  const int row = (int)cr_div((uint32_t)r, g.mRW, g.sRW), col = r - row * g.RW;
  n = (int)cr_div(S, g.mTS, g.sTS);
  t = (int)(S - (uint32_t)n * (uint32_t)g.TS) - g.ot;
  h = row - g.oh; w = col - g.ow;
  return (unsigned)t < (unsigned)g.T && (unsigned)h < (unsigned)g.H && (unsigned)w < (unsigned)g.W && n < g.N;
}
__device__ __forceinline__ int64_t cr_flat(const UcGeom& g, int n, int t, int h, int w) {
  return (((int64_t)n * g.TS + t + g.ot) * g.HP + h + g.oh) * g.RW + w + g.ow;
}
__device__ __forceinline__ uint4 cr_pack8_relu(const float* v) {
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(ow[e]) : "f"(v[2 * e + 1]), "f"(v[2 * e]));
  return o;
}

__global__ void __launch_bounds__(CR_THREADS, 1) conv_ring_kernel(const __grid_constant__ ConvRingP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[CR_NST], empty_bar[CR_NST], wfull[2], wempty[2], tfull[2][4], tempty[2][4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rb = (uint32_t)p.units * 16u;                       // bytes of one plane's region
  const uint32_t slot_bytes = 8u * rb, stage_bytes = 2u * slot_bytes;
  uint8_t* const wsm = smem;                                          // weight ring: 2 x CR_WCHUNK
  uint8_t* const ones = smem + 2 * CR_WCHUNK;
  uint8_t* const biasb = ones + 2048;
  uint8_t* const stages = biasb + 2048;
  SrStep* const tab = reinterpret_cast<SrStep*>(stages + (size_t)p.nst * stage_bytes);
  {
    const uint2* src = reinterpret_cast<const uint2*>(p.steps + (size_t)(2 * blockIdx.x) * (size_t)p.nsteps);
    for (int i = tid; i < 2 * p.nsteps; i += CR_THREADS) reinterpret_cast<uint2*>(tab)[i] = src[i];
  }
  if (tid == 0) {
    for (int i = 0; i < CR_NST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 2); }
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < 4; ++i) { mbar_init(&tfull[s][i], 1); mbar_init(&tempty[s][i], 4); }
    fence_barrier_init();
  }
  if (warp == 3) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid < 128) {
    reinterpret_cast<uint4*>(ones)[tid] = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 64 && p.bias) {
      const float b = p.bias[tid];
      const __nv_bfloat16 hi = __float2bfloat16_rn(b);
      const float e1 = b - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(e1);
      const __nv_bfloat16 lo = __float2bfloat16_rn(e1 - __bfloat162float(mid));
      o.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
      o.y = (uint32_t)__bfloat16_as_ushort(lo);
    }
    reinterpret_cast<uint4*>(biasb)[tid] = o;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const SrStep* const st0 = tab;
  unsigned long long gt0 = 0;
  long long ck0 = 0;
  if (p.dbg && tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0)); ck0 = clock64(); }

  if (warp == 0) {
    // ------------------------------------------------ activation producer: per step the 8 planes of both slots' regions
    const int slot = (lane >> 3) & 1, plane = lane & 7;
    const SrStep* const st = st0 + (size_t)slot * (size_t)p.nsteps;
    const uint32_t dst0 = smem_u32(stages) + (uint32_t)slot * slot_bytes + (uint32_t)plane * rb;
    const __nv_bfloat16* const xp = p.x + (int64_t)plane * p.x_plane_stride;
    for (int k = 0; k < p.nsteps; ++k) {
      const int stage = k % p.nst;
      const uint32_t ph = (uint32_t)(k / p.nst) & 1u;
      const SrStep sd = st[k];
      const bool act = lane < 16 && (sd.flags & SR_ACTIVE);
      const uint32_t total = __reduce_add_sync(0xffffffffu, act ? rb : 0u);
      mbar_wait(&empty_bar[stage], ph ^ 1u);
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], total);
      __syncwarp();
      if (act) bulk_s2(dst0 + (uint32_t)stage * stage_bytes, xp + ((int64_t)sd.in_pos + (int64_t)p.start) * 8, rb, &full_bar[stage]);
      __syncwarp();
    }
  } else if (warp == 3) {
    // ------------------------------------------------ weight producer: every step streams the four K16 chunks through the 2-stage ring
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
      int i = 0;
      for (int k = 0; k < p.nsteps; ++k)
        for (int c = 0; c < CR_K16; ++c, ++i) {
          const int ws = i & 1;
          mbar_wait(&wempty[ws], ((uint32_t)(i >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&wfull[ws], (uint32_t)CR_WCHUNK);
          bulk_s2(smem_u32(wsm) + (uint32_t)ws * (uint32_t)CR_WCHUNK, wsrc + (size_t)c * CR_WCHUNK, (uint32_t)CR_WCHUNK, &wfull[ws]);
        }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------ MMA issuer of slot (warp - 1): whole warp runs the loop, one lane issues
    const int slot = __shfl_sync(0xffffffffu, warp, 0) - 1;
    const SrStep* const st = st0 + (size_t)slot * (size_t)p.nsteps;
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint64_t desc_hi64 = ((uint64_t)8 << 32) | ((uint64_t)1 << 46);          // SBO = 128 B, descriptor version 1
    const uint32_t id192 = idesc_bf16(128, 192), id128 = idesc_bf16(128, 128), id64 = idesc_bf16(128, 64);
    const uint32_t tb = tmem + (uint32_t)slot * 256u;
    const uint64_t d_ones = desc_hi64 | (uint64_t)(smem_u32(ones) >> 4);                           // LBO = 0
    const uint64_t d_bias = desc_hi64 | (uint64_t)((smem_u32(biasb) >> 4) | (64u << 16));         // LBO = 64 columns x 16 B
    const uint32_t w_ring = (smem_u32(wsm) >> 4) | (192u << 16);                                   // LBO = 192 columns x 16 B
    const uint32_t a_slot = ((smem_u32(stages) + (uint32_t)slot * slot_bytes) >> 4) | ((uint32_t)p.units << 16);   // LBO = one plane's region
    const uint32_t rw = (uint32_t)p.g.RW, ru = (uint32_t)p.units;
    int q = 0, wi = 0;
    for (int k = 0; k < p.nsteps; ++k) {
      const int stage = k % p.nst;
      const uint32_t ph = (uint32_t)(k / p.nst) & 1u;
      const uint32_t flags = __shfl_sync(0xffffffffu, st[k].flags, 0);
      mbar_wait(&full_bar[stage], ph);
      tc_fence_after();
      const bool active = (flags & SR_ACTIVE) != 0;
      uint32_t c = 0;
      if (active) {
        if (q >= 2) { mbar_wait(&tempty[slot][(q - 2) & 3], (uint32_t)((q - 2) >> 2) & 1u); tc_fence_after(); }
        c = (uint32_t)(-q) & 3u;              // first block of this step's D region: [out(t+1) | out(t) | out(t-1)]
        if (flags & SR_FIRST) {
          mma_bf16_ss_pred(tb + c * 64u, d_ones, d_bias, id64, 0u, leader);
          mma_bf16_ss_pred(tb + ((c + 1u) & 3u) * 64u, d_ones, d_bias, id64, 0u, leader);
          mma_bf16_ss_pred(tb + ((c + 2u) & 3u) * 64u, d_ones, d_bias, id64, 0u, leader);
        } else {
          mma_bf16_ss_pred(tb + c * 64u, d_ones, d_bias, id64, 0u, leader);
        }
      }
      const uint32_t ab = a_slot + (((uint32_t)stage * stage_bytes) >> 4);
#pragma unroll 1
      for (int kc = 0; kc < CR_K16; ++kc, ++wi) {
        const int ws = wi & 1;
        mbar_wait(&wfull[ws], (uint32_t)(wi >> 1) & 1u);
        tc_fence_after();
        if (active && !((p.skip & 4) && kc >= 1)) {
          uint32_t wj = w_ring + (uint32_t)ws * (uint32_t)(CR_WCHUNK >> 4);
          const uint32_t ac = ab + (uint32_t)kc * 2u * ru;          // planes 2kc, 2kc+1 of the slot's stage
#pragma unroll 1
          for (int dh = 0; dh < 3; ++dh) {
            uint32_t at = ac + (uint32_t)dh * rw;
#pragma unroll 1
            for (int dw = 0; dw < 3; ++dw, ++at, wj += (uint32_t)((2 * 192 * 16) >> 4)) {
              const uint64_t da = desc_hi64 | (uint64_t)at, db = desc_hi64 | (uint64_t)wj;
              if (c <= 1u) {
                mma_bf16_ss_pred(tb + c * 64u, da, db, id192, 1u, leader);
              } else if (c == 2u) {          // blocks 2, 3, then (wrapped) block 0
                mma_bf16_ss_pred(tb + 128u, da, db, id128, 1u, leader);
                mma_bf16_ss_pred(tb, da, db + 128u, id64, 1u, leader);
              } else {                       // block 3, then (wrapped) blocks 0, 1
                mma_bf16_ss_pred(tb + 192u, da, db, id64, 1u, leader);
                mma_bf16_ss_pred(tb, da, db + 64u, id128, 1u, leader);
              }
            }
          }
        }
        mma_commit_pred(&wempty[ws], leader);            // this issuer is done with the weight stage
      }
      if (active) {
        mma_commit_pred(&tfull[slot][q & 3], leader);    // out(t-1) of this slot is complete
        ++q;
      }
      mma_commit_pred(&empty_bar[stage], leader);        // the activation stage is free once the MMAs that read it have completed
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue: drain the completed block (thread = position)
    const int slot = (warp - 4) >> 2, quarter = warp & 3;
    const SrStep* const st = st0 + (size_t)slot * (size_t)p.nsteps;
    const int i = quarter * 32 + lane;
    const bool has_res = p.res != nullptr;
    int q = 0;
    for (int k = 0; k < p.nsteps; ++k) {
      const SrStep sd = st[k];
      if (!(sd.flags & SR_ACTIVE)) continue;
      mbar_wait(&tfull[slot][q & 3], (uint32_t)(q >> 2) & 1u);
      tc_fence_after();
      if ((sd.flags & SR_STORE) && !(p.skip & 1)) {
        const uint32_t blk = (((uint32_t)(-q) & 3u) + 2u) & 3u;
        const uint32_t ta = tmem + (uint32_t)slot * 256u + blk * 64u + ((uint32_t)(quarter * 32) << 16);
        const bool inchunk = i < (int)((sd.flags >> 8) & 0xffu);
        const int64_t P = (int64_t)sd.in_pos - p.g.SL + i;
        int n = 0, t = 0, h = 0, w = 0;
        const bool valid = inchunk && cr_decode(p.g, P, n, t, h, w);
        const uint32_t vmask = valid ? 0xffffffffu : 0u;
        int64_t dst = P * 8;
        bool store = inchunk;                                   // plain: pad positions of the slab are written as zeros
        if (p.y_mode == UC_Y_PARITY) {
          store = valid;
          if (valid) dst = (int64_t)((h & 1) * 2 + (w & 1)) * p.y_set_stride + cr_flat(p.g2, n, t, h >> 1, w >> 1) * 8;
        }
        char* yp = reinterpret_cast<char*>(p.y + dst);
        const char* rp = reinterpret_cast<const char*>(p.res + P * 8);
        const int64_t yps = p.y_plane_stride * 2, rps = p.res_plane_stride * 2;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32, yp += 4 * yps, rp += 4 * rps) {
          float v[32];
          uint4 rr[4];
          tmem_ld32(ta + (uint32_t)c0, v);
          if (has_res) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) rr[qq] = valid ? *reinterpret_cast<const uint4*>(rp + qq * rps) : make_uint4(0, 0, 0, 0);
          }
          tmem_ld_wait();
          if (has_res) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const __nv_bfloat162* rb2 = reinterpret_cast<const __nv_bfloat162*>(&rr[qq]);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(rb2[e]); v[8 * qq + 2 * e] += f.x; v[8 * qq + 2 * e + 1] += f.y; }
            }
          }
          if (store) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              uint4 o = cr_pack8_relu(v + 8 * qq);
              o.x &= vmask; o.y &= vmask; o.z &= vmask; o.w &= vmask;
              *reinterpret_cast<uint4*>(yp + qq * yps) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[slot][q & 3]);
      ++q;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && tid == 0 && blockIdx.x < 160) {
    unsigned long long gt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    long long* d = p.dbg + 4 * blockIdx.x;
    d[0] = (long long)gt0; d[1] = (long long)gt1; d[2] = ck0; d[3] = clock64();
  }
  if (warp == 3) tmem_dealloc(tmem, 512);
}

}  // namespace

size_t conv_ring_smem_bytes(const ConvRingP& p) {
  return (size_t)2 * CR_WCHUNK + 4096 + (size_t)p.nst * 2 * 8 * ((size_t)p.units * 16) + (size_t)2 * p.nsteps * sizeof(SrStep) + 1024;
}

cudaError_t conv_ring_device_init() {
  return cudaFuncSetAttribute(conv_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 223 * 1024);
}

void launch_conv_ring(const ConvRingP& p, int grid, cudaStream_t s) {
  if (grid <= 0 || p.nsteps <= 0) return;
  conv_ring_kernel<<<grid, CR_THREADS, conv_ring_smem_bytes(p), s>>>(p);
  count_launch();
}

}  // namespace lsd
