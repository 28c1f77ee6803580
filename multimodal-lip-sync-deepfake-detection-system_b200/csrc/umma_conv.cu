// tcgen05 "flat shift-GEMM" convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// Replaces nn.Conv3d / Conv2d / Conv1d / Linear + eval BatchNorm + ReLU/GELU + residual of the reference
// (app/models/visual_encoder.py:46-87, audio_encoder.py:34-89, artifact_detector.py:74-93,41-43, temporal.py:35-51)
// on the bf16 path.  See umma_conv.cuh for the padded planar activation layout that turns every tap into a shift.
//
// One CTA = MT consecutive 128-position M tiles x all Cout (<= 256) columns; accumulators in TMEM (MT*Cout columns).
// Pipeline (mbarrier ring, `stages` deep), per (k16 channel chunk, band):
//   producer thread : 3 bulk async copies (TMA engine, 1-D): the band's two 8-channel planes + the band's packed weights
//   MMA thread      : ntaps x MT tcgen05.mma (M=128, N=Cout, K=16); the A descriptor of a tap is the band base shifted by the
//                     tap offset (K-major, SWIZZLE_NONE: rows 16 B apart, SBO = 128 B, LBO = band length * 16 B)
//   tcgen05.commit releases the stage; the last commit signals the epilogue.
// Epilogue (all 4 warps): tcgen05.ld -> +bias (+residual) -> act -> zero pads -> bf16 -> 16 B coalesced stores per plane.
#include "umma_conv.cuh"
#include "token_kernels.cuh"

#include "lsd_kernels.h"
#include "umma.cuh"

namespace lsd {

using namespace umma;

__device__ __forceinline__ float uc_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  return v;
}

// (P_total < 2^31 is enforced on the host: 32-bit divisions only)
__device__ __forceinline__ bool uc_decode(const UcGeom& g, int64_t P, int& n, int& t, int& h, int& w) {
  if (P < 0 || P >= g.P_total) return false;
  const uint32_t Pu = (uint32_t)P;
  const uint32_t S = Pu / (uint32_t)g.SL;
  const int r = (int)(Pu - S * (uint32_t)g.SL);
  const int row = r / g.RW, col = r - row * g.RW;
  n = (int)(S / (uint32_t)g.TS);
  t = (int)(S - (uint32_t)n * (uint32_t)g.TS) - g.ot;
  h = row - g.oh; w = col - g.ow;
  return (unsigned)t < (unsigned)g.T && (unsigned)h < (unsigned)g.H && (unsigned)w < (unsigned)g.W && n < g.N;
}

__device__ __forceinline__ int64_t uc_flat(const UcGeom& g, int n, int t, int h, int w) {
  return (((int64_t)n * g.TS + t + g.ot) * g.HP + h + g.oh) * g.RW + w + g.ow;
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 o;
  __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  return o;
}
__device__ __forceinline__ void unpack8(const uint4& r, float* v) {
  const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(rb[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}

// Persistent, warp-specialised (14 warps):
//   warps 0-3  bulk-copy producers (measured on B200: one warp sustains only ~1 cp.async.bulk per ~100-130 cycles whatever the
//              number of active lanes, and the rate scales with the number of issuing warps -> the copies of a stage are
//              dealt round-robin to four warps),
//   warps 4-5  MMA issuers (each owns half of the tile's M-tiles),
//   warps 6-13 epilogue (two warps per TMEM lane quarter, each handling every other 32-column chunk).
// Each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the accumulators are double-buffered in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1, and the smem ring keeps streaming across tile boundaries.
constexpr int UC_THREADS = 448;
constexpr int UC_PROD_WARPS = 4, UC_MMA_WARP0 = 4, UC_EPI_WARP0 = 6, UC_EPI_WARPS = 8;

// GENERIC = false: lean epilogue of the convolution layers (bias, ReLU/none, optional bf16 residual, planar / parity-split bf16
// store).  GENERIC = true: everything (fp32 rows in/out, split-bf16 hi/lo outputs and residuals, GELU) for the audio encoder and
// the token-path GEMMs.  Two instantiations keep each one small enough for the instruction cache.
template <bool GENERIC>
__global__ void __launch_bounds__(UC_THREADS, 1) umma_conv_kernel(const __grid_constant__ UmmaConvP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  if (dbg && tid == 0) p.dbg[0] = clock64();
  const int S = p.MT * 128;
  const uint32_t stage_bytes = p.a_stage_bytes + p.w_stage_bytes;
  const int num_tiles = (int)((p.g.P_total + S - 1) / S);
  const int slice = blockIdx.y, ch0 = slice * p.Cout;
  const uint32_t buf_cols = (uint32_t)(p.MT * p.Cout);

  const int n_issuers = p.MT > 1 ? 2 : 1;   // MMA-issuing warps (M-tiles are independent accumulators)
  if (tid == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], n_issuers); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], n_issuers); mbar_init(&tempty_bar[i], UC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == UC_EPI_WARP0) { tmem_alloc(&tmem_base_s, p.tmem_cols); tmem_relinquish(); }
  for (int i = tid; i < p.Cout; i += UC_THREADS) bias_s[i] = p.bias ? p.bias[ch0 + i] : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (dbg && tid == 0) p.dbg[1] = clock64();   // prologue done

  if (warp < UC_PROD_WARPS) {
    // ------------------------------------------------ producers (every warp runs the loop; lane 0 of each issues its share)
    int stage = 0, dbg_it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int64_t P0 = (int64_t)tile * S;
      for (int gi = 0; gi < p.ngroups; ++gi) {
        const UcGroup& g = p.groups[gi];
        for (int c0 = 0; c0 < g.k16; c0 += p.kpack) {
          const int nc = min(p.kpack, g.k16 - c0);
          for (int b = g.band_begin; b < g.band_end; ++b) {
            const UcBand& bd = p.bands[b];
            mbar_wait(&empty_bar[stage], ph ^ 1u);
            if (dbg && warp == 0 && lane == 0 && dbg_it < 24 && tile == (int)blockIdx.x) p.dbg[32 + dbg_it++] = clock64();
            if (lane == 0) {
              const uint32_t bytesW = (uint32_t)bd.ntaps * (uint32_t)p.Cout * 32u;
              uint8_t* sa = smem + (size_t)stage * stage_bytes;
              uint8_t* sw = sa + p.a_stage_bytes;
              const __nv_bfloat16* src = bd.base + (int64_t)c0 * bd.chunk_stride + (P0 + bd.start) * 8;
              const uint32_t bytesA = bd.toeplitz ? (uint32_t)(S + bd.len_extra + 1 + 2 * (nc - 1)) * 16u : (uint32_t)(S + bd.len_extra) * 16u;
              const int n_a = bd.toeplitz ? 1 : 2 * nc;   // Toeplitz: the K chunks are 32-byte shifts of one row region
              // single-band groups (Linear layers): the weights of consecutive K chunks are contiguous -> one copy
              const bool w_merged = (g.band_end - g.band_begin) == 1;
              const int n_w = w_merged ? 1 : nc;
              if (warp == 0) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)n_a * bytesA + (uint32_t)nc * bytesW);
              for (int idx = warp; idx < n_a + n_w; idx += UC_PROD_WARPS) {
                if (idx < n_a) {
                  const int j = idx >> 1, pl = idx & 1;     // K chunk j, plane pl (planar); idx == 0 only for Toeplitz
                  bulk_g2s(sa + (size_t)idx * bytesA, src + (int64_t)j * bd.chunk_stride + (int64_t)pl * bd.plane_stride, bytesA, &full_bar[stage]);
                } else {
                  const int j = idx - n_a;
                  const __nv_bfloat16* wsrc = p.w + g.w_off + (int64_t)slice * g.slice_stride +
                                              ((int64_t)(c0 + j) * g.taps_total + bd.tap_begin) * (int64_t)p.Cout * 16;
                  bulk_g2s(sw + (size_t)j * bytesW, wsrc, w_merged ? (uint32_t)nc * bytesW : bytesW, &full_bar[stage]);
                }
              }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == UC_MMA_WARP0 || warp == UC_MMA_WARP0 + 1) {
    // ------------------------------------------------ MMA issuers (whole warp runs the loop; one elected lane issues)
    if (warp == UC_MMA_WARP0 + 1 && n_issuers == 1) goto done;
    const int mt_lo = (warp == UC_MMA_WARP0) ? 0 : p.MT / 2;                         // this warp's M-tiles: [mt_lo, mt_lo + mt_n)
    const int mt_n = n_issuers == 1 ? p.MT : p.MT / 2;
    // One thread feeds the tensor core: keep the per-instruction work to a few 32-bit adds.  A descriptor is
    // (constant high part) | (start address >> 4); tap / M-tile / K-chunk offsets are added in 16-byte units.
    const uint32_t idesc = idesc_bf16(128, p.Cout);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t desc_hi = ((uint64_t)8 << 32) | ((uint64_t)1 << 46);            // SBO = 128 B, descriptor version 1
    const uint64_t db_hi = desc_hi | ((uint64_t)(uint32_t)p.Cout << 16);           // LBO(B) = Cout * 16 B
    const uint32_t tap_w = (uint32_t)p.Cout * 2u;                                  // Cout * 32 B per tap, in 16 B units
    int stage = 0, lt = 0, dbg_it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int buf = p.nbuf == 2 ? (lt & 1) : 0;
      const uint32_t use = p.nbuf == 2 ? ((uint32_t)lt >> 1) : (uint32_t)lt;   // how many times this buffer was used before
      const uint32_t tb = tmem_base + (uint32_t)buf * buf_cols + (uint32_t)(mt_lo * p.Cout);
      mbar_wait(&tempty_bar[buf], (use & 1u) ^ 1u);   // epilogue has drained this accumulator buffer
      tc_fence_after();
      uint32_t acc = 0;
      for (int gi = 0; gi < p.ngroups; ++gi) {
        const UcGroup& g = p.groups[gi];
        for (int c0 = 0; c0 < g.k16; c0 += p.kpack) {
          const int nc = min(p.kpack, g.k16 - c0);
          for (int b = g.band_begin; b < g.band_end; ++b) {
            const UcBand& bd = p.bands[b];
            const int ntaps = bd.ntaps;
            const uint32_t unitsA = (uint32_t)(S + bd.len_extra);
            const uint64_t da_hi = desc_hi | ((uint64_t)(bd.toeplitz ? 1u : unitsA) << 16);
            const uint32_t a_chunk = bd.toeplitz ? 2u : 2u * unitsA;
            const uint32_t w_chunk = (uint32_t)ntaps * tap_w;
            mbar_wait(&full_bar[stage], ph);
            tc_fence_after();
            if (dbg && warp == UC_MMA_WARP0 && lane == 0 && acc == 0 && lt == 0) p.dbg[2] = clock64();   // first stage landed
            if (dbg && warp == UC_MMA_WARP0 && lane == 0 && lt == 0 && dbg_it < 24) p.dbg[8 + dbg_it++] = clock64();
            const uint32_t sa = (smem_base + (uint32_t)stage * stage_bytes) >> 4;
            const uint32_t sw = sa + (p.a_stage_bytes >> 4);
            if (elect_one()) {
              uint32_t accl = acc;   // 0 only for the very first MMA of each accumulator
              for (int j = 0; j < nc; ++j) {
                const uint32_t aj = sa + (uint32_t)j * a_chunk, wj = sw + (uint32_t)j * w_chunk;
#pragma unroll 1
                for (int tp = 0; tp < ntaps; ++tp) {   // (not unrolled: the kernel must stay inside the instruction cache)
                  const uint64_t db = db_hi | (uint64_t)(wj + (uint32_t)tp * tap_w);
                  const uint32_t at = aj + (uint32_t)bd.rel[tp] + (uint32_t)mt_lo * 128u;
                  mma_bf16_ss(tb, da_hi | (uint64_t)at, db, idesc, accl);
                  if (mt_n > 1) mma_bf16_ss(tb + (uint32_t)p.Cout, da_hi | (uint64_t)(at + 128u), db, idesc, accl);
                  accl = 1u;
                }
              }
              mma_commit(&empty_bar[stage]);  // frees the stage once the MMAs that read it have completed
            }
            acc = 1u;
            __syncwarp();
            if (++stage == p.stages) { stage = 0; ph ^= 1u; }
          }
        }
      }
      if (elect_one()) mma_commit(&tfull_bar[buf]);   // accumulator of this tile complete -> epilogue
      __syncwarp();
      if (dbg && warp == UC_MMA_WARP0 && lane == 0 && lt == 0) p.dbg[3] = clock64();   // all MMAs of the first tile issued
    }
  } else {
    // ------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4; thread == tile row)
    const int quarter = warp & 3;
    const int half = (warp - UC_EPI_WARP0) >> 2;                 // which 32-column chunks this warp handles (even / odd)
    const float act_lo = p.act == ACT_RELU ? 0.0f : -INFINITY;   // ReLU as one FMNMX; GELU (token GEMMs only) branches once per chunk
    const int rowt = quarter * 32 + lane;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int buf = p.nbuf == 2 ? (lt & 1) : 0;
      const uint32_t use = p.nbuf == 2 ? ((uint32_t)lt >> 1) : (uint32_t)lt;
      const uint32_t tb = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(quarter * 32) << 16);
      const int64_t P0 = (int64_t)tile * S;
      mbar_wait(&tfull_bar[buf], use & 1u);
      tc_fence_after();
      if (dbg && warp == UC_EPI_WARP0 && lane == 0 && lt == 0) p.dbg[4] = clock64();   // first accumulator complete
      for (int m = 0; m < p.MT; ++m) {
        const int64_t P = P0 + (int64_t)m * 128 + rowt;
        int n, t, h, w;
        const bool valid = uc_decode(p.g, P, n, t, h, w);
        const bool inrange = P < p.g.P_total;
        const int64_t outer = valid ? ((int64_t)n * p.g.T + t) * p.g.H + h : 0;
        int64_t dst = P * 8;
        if (valid && p.y_mode == UC_Y_PARITY)
          dst = (int64_t)((h & 1) * 2 + (w & 1)) * p.y_set_stride + uc_flat(p.g2, n, t, h >> 1, w >> 1) * 8;
        else if (valid && p.y_mode == UC_Y_PARITY_H)
          dst = (int64_t)((h & 1) * 2) * p.y_set_stride + uc_flat(p.g2, n, t, h >> 1, w) * 8;
        const bool store_planar = (p.y_mode == UC_Y_PLAIN && inrange) || (p.y_mode >= UC_Y_PARITY && valid);
        // this warp's columns: [half*Cout/2, (half+1)*Cout/2)
        const int cbeg = half * (p.Cout >> 1), cend = cbeg + (p.Cout >> 1);
        if constexpr (!GENERIC) {
          // lean path: 8 columns (one 16-byte plane entry) per step, pointers advanced by one plane per step
          __nv_bfloat16* yp = p.y + (int64_t)((ch0 + cbeg) >> 3) * p.y_plane_stride + dst;
          const __nv_bfloat16* rp = p.res + (int64_t)((ch0 + cbeg) >> 3) * p.res_plane_stride + P * 8;
          const bool has_res = p.res != nullptr && valid;
          const float* bp = &bias_s[cbeg];
          uint32_t ta = tb + (uint32_t)(m * p.Cout + cbeg);
#pragma unroll 1
          for (int c = cbeg; c < cend; c += 8, yp += p.y_plane_stride, rp += p.res_plane_stride, bp += 8, ta += 8) {
            float v[8];
            tmem_ld8(ta, v);
            uint4 rr = make_uint4(0, 0, 0, 0);
            if (has_res) rr = *reinterpret_cast<const uint4*>(rp);   // overlaps the TMEM load
            const float4 b0 = *reinterpret_cast<const float4*>(bp);
            const float4 b1 = *reinterpret_cast<const float4*>(bp + 4);
            tmem_ld_wait();
            float f[8];
            unpack8(rr, f);
            v[0] = fmaxf(v[0] + b0.x + f[0], act_lo); v[1] = fmaxf(v[1] + b0.y + f[1], act_lo);
            v[2] = fmaxf(v[2] + b0.z + f[2], act_lo); v[3] = fmaxf(v[3] + b0.w + f[3], act_lo);
            v[4] = fmaxf(v[4] + b1.x + f[4], act_lo); v[5] = fmaxf(v[5] + b1.y + f[5], act_lo);
            v[6] = fmaxf(v[6] + b1.z + f[6], act_lo); v[7] = fmaxf(v[7] + b1.w + f[7], act_lo);
            if (store_planar) *reinterpret_cast<uint4*>(yp) = valid ? pack8(v) : make_uint4(0, 0, 0, 0);
          }
        } else {
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          const int nq = min(4, (cend - c0) >> 3);
          // residual prefetch for the whole chunk: the global-load latency is paid once, not once per 8-column step
          uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0;
          if (valid && p.res) {
            const __nv_bfloat16* rp = p.res + (int64_t)((ch0 + c0) >> 3) * p.res_plane_stride + P * 8;
            r0 = *reinterpret_cast<const uint4*>(rp);
            if (nq > 1) r1 = *reinterpret_cast<const uint4*>(rp + p.res_plane_stride);
            if (nq > 2) r2 = *reinterpret_cast<const uint4*>(rp + 2 * p.res_plane_stride);
            if (nq > 3) r3 = *reinterpret_cast<const uint4*>(rp + 3 * p.res_plane_stride);
          }
#pragma unroll 1
          for (int q = 0; q < nq; ++q) {   // 8 columns per step keeps the epilogue code small (instruction cache)
            const int c = c0 + 8 * q;
            float v[8];
            tmem_ld8(tb + (uint32_t)(m * p.Cout + c), v);
            tmem_ld_wait();
            if (valid) {
              const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[c]);
              const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[c + 4]);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              if (p.res) {
                float f[8];
                const uint4 rr = q == 0 ? r0 : (q == 1 ? r1 : (q == 2 ? r2 : r3));
                unpack8(rr, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += f[e];
                if (p.res_lo) {
                  unpack8(*reinterpret_cast<const uint4*>(p.res_lo + (int64_t)((ch0 + c) >> 3) * p.res_plane_stride + P * 8), f);
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] += f[e];
                }
              }
              if (p.res32) {
                const float4* r4 = reinterpret_cast<const float4*>(p.res32 + (outer * p.g.W + w) * p.res32_ld + ch0 + c);
                const float4 q0 = r4[0], q1 = r4[1];
                v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w; v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
              }
              if (p.act == ACT_GELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.5f * v[e] * (1.0f + erff(v[e] * 0.70710678118654752f));
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], act_lo);
              }
              if (p.y32) {
                float4* o = reinterpret_cast<float4*>(p.y32 + (outer * p.y32_outer_stride + w + p.y32_row_off) * p.y32_ld + ch0 + c);
                o[0] = make_float4(v[0], v[1], v[2], v[3]);
                o[1] = make_float4(v[4], v[5], v[6], v[7]);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.0f;
            }
            if (store_planar) {
              const uint4 hi4 = pack8(v);
              const int64_t off = (int64_t)((ch0 + c) >> 3) * p.y_plane_stride + dst;
              *reinterpret_cast<uint4*>(p.y + off) = hi4;
              if (p.ylo) {
                float hf[8];
                unpack8(hi4, hf);
#pragma unroll
                for (int e = 0; e < 8; ++e) hf[e] = v[e] - hf[e];
                *reinterpret_cast<uint4*>(p.ylo + off) = pack8(hf);
              }
            }
          }
        }
        }
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (dbg && warp == UC_EPI_WARP0 && lane == 0 && lt == 0) p.dbg[5] = clock64();   // first epilogue done
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == UC_EPI_WARP0) tmem_dealloc(tmem_base, p.tmem_cols);
  if (dbg && tid == 0) { p.dbg[6] = clock64(); p.dbg[7] = num_tiles; }
}

size_t umma_conv_smem_bytes(const UmmaConvP& p) { return (size_t)p.stages * (p.a_stage_bytes + p.w_stage_bytes) + 1024; }

void launch_umma_conv(const UmmaConvP& p, int n_slices, cudaStream_t s, int max_ctas) {
  static bool attr_set = false;
  if (!attr_set) {
    // the opt-in limit (227 KB) covers static + dynamic shared memory; ~1.3 KB is static (barriers, bias)
    cudaFuncSetAttribute(umma_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(umma_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    attr_set = true;
  }
  const int S = p.MT * 128;
  const int tiles = (int)((p.g.P_total + S - 1) / S);
  static int num_sms = 0;
  if (num_sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
  // persistent grid: one CTA per SM (per Cout slice), each walking tiles with stride gridDim.x
  const int budget = (max_ctas > 0 && max_ctas < num_sms) ? max_ctas : num_sms;   // side-stream launches leave SMs to the main stream
  int gx = (budget + n_slices - 1) / n_slices;
  gx = gx < 1 ? 1 : (gx > tiles ? tiles : gx);
  const bool generic = p.y32 || p.res32 || p.ylo || p.res_lo || p.act == ACT_GELU || p.y_mode == UC_Y_NONE;
  if (generic) umma_conv_kernel<true><<<dim3(gx, n_slices), UC_THREADS, umma_conv_smem_bytes(p), s>>>(p);
  else umma_conv_kernel<false><<<dim3(gx, n_slices), UC_THREADS, umma_conv_smem_bytes(p), s>>>(p);
  count_launch();
}

// ================================================================================================
// planar-layout glue kernels (memory-bound, 16-byte accesses, one thread per (position, 8-channel chunk))
// ================================================================================================
__global__ void pack_planar_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t plane_stride, int64_t set_stride,
                                   UcGeom g, UcGeom g2, int C, int parity, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // position-major within a chunk so that a warp writes 512 contiguous bytes of one plane
  const int64_t npos = (int64_t)g.N * g.T * g.H * g.W;
  const int chunk = (int)(i / npos);
  int64_t r = i - (int64_t)chunk * npos;
  const int w = (int)(r % g.W); r /= g.W;
  const int h = (int)(r % g.H); r /= g.H;
  const int t = (int)(r % g.T);
  const int n = (int)(r / g.T);
  const float* src = x + ((((int64_t)n * g.T + t) * g.H + h) * g.W + w) * C + chunk * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = src[e];
  int64_t dst;
  if (parity) dst = (int64_t)((h & 1) * 2 + (w & 1)) * set_stride + uc_flat(g2, n, t, h >> 1, w >> 1) * 8;
  else dst = uc_flat(g, n, t, h, w) * 8;
  *reinterpret_cast<uint4*>(y + (int64_t)chunk * plane_stride + dst) = pack8(v);
}
void launch_pack_planar(const float* x, __nv_bfloat16* y, int64_t plane_stride, int64_t set_stride, UcGeom g, UcGeom g2, int C,
                        int parity, cudaStream_t s) {
  const int64_t total = (int64_t)g.N * g.T * g.H * g.W * (C / 8);
  if (total == 0) return;
  pack_planar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, plane_stride, set_stride, g, g2, C, parity, total);
  count_launch();
}

__global__ void unpack_planar_kernel(const __nv_bfloat16* __restrict__ x, int64_t plane_stride, UcGeom g, int C, float* __restrict__ y,
                                     int ld, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int nch = C / 8;
  const int chunk = (int)(i % nch);
  int64_t r = i / nch;
  const int64_t row = r;
  const int w = (int)(r % g.W); r /= g.W;
  const int h = (int)(r % g.H); r /= g.H;
  const int t = (int)(r % g.T);
  const int n = (int)(r / g.T);
  const uint4 v = *reinterpret_cast<const uint4*>(x + (int64_t)chunk * plane_stride + uc_flat(g, n, t, h, w) * 8);
  float f[8];
  unpack8(v, f);
  float4* o = reinterpret_cast<float4*>(y + row * ld + chunk * 8);
  o[0] = make_float4(f[0], f[1], f[2], f[3]);
  o[1] = make_float4(f[4], f[5], f[6], f[7]);
}
void launch_unpack_planar(const __nv_bfloat16* x, int64_t plane_stride, UcGeom g, int C, float* y, int ld, cudaStream_t s) {
  const int64_t total = (int64_t)g.N * g.T * g.H * g.W * (C / 8);
  if (total == 0) return;
  unpack_planar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, plane_stride, g, C, y, ld, total);
  count_launch();
}

// Deterministic mean: one block per (row, chunk); threads stride over the row's positions, fixed-order tree in shared memory.
__global__ void planar_mean_kernel(const __nv_bfloat16* __restrict__ x, int64_t plane_stride, UcGeom g, float* __restrict__ y, int ld,
                                   int per_window) {
  __shared__ float part[256][9];
  const int row = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
  const int count = per_window ? g.T * g.H * g.W : g.H * g.W;
  const int n = per_window ? row : row / g.T;
  const int t0 = per_window ? 0 : row % g.T;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = tid; i < count; i += blockDim.x) {
    const int w = i % g.W;
    const int r = i / g.W;
    const int h = r % g.H;
    const int t = t0 + r / g.H;
    const uint4 v = *reinterpret_cast<const uint4*>(x + (int64_t)chunk * plane_stride + uc_flat(g, n, t, h, w) * 8);
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[tid][e] = acc[e];
  __syncthreads();
  for (int sft = blockDim.x >> 1; sft > 0; sft >>= 1) {
    if (tid < sft) {
#pragma unroll
      for (int e = 0; e < 8; ++e) part[tid][e] += part[tid + sft][e];
    }
    __syncthreads();
  }
  if (tid < 8) y[(int64_t)row * ld + chunk * 8 + tid] = part[0][tid] / (float)count;
}
void launch_planar_mean(const __nv_bfloat16* x, int64_t plane_stride, UcGeom g, int C, float* y, int ld, int per_window, cudaStream_t s) {
  const int rows = per_window ? g.N : g.N * g.T;
  if (rows == 0) return;
  const int count = per_window ? g.T * g.H * g.W : g.H * g.W;
  int threads = 32;
  while (threads < 256 && threads < count) threads <<= 1;
  planar_mean_kernel<<<dim3(rows, C / 8), threads, 0, s>>>(x, plane_stride, g, y, ld, per_window);
  count_launch();
}

__global__ void planar_delta_kernel(const __nv_bfloat16* __restrict__ x, int64_t xs, UcGeom g, __nv_bfloat16* __restrict__ d, int64_t ds,
                                    UcGeom gd, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t npos = (int64_t)gd.N * gd.T * gd.H * gd.W;
  const int chunk = (int)(i / npos);
  int64_t r = i - (int64_t)chunk * npos;
  const int w = (int)(r % gd.W); r /= gd.W;
  const int h = (int)(r % gd.H); r /= gd.H;
  const int t = (int)(r % gd.T);
  const int n = (int)(r / gd.T);
  float a[8], b[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(x + (int64_t)chunk * xs + uc_flat(g, n, t + 1, h, w) * 8), a);
  unpack8(*reinterpret_cast<const uint4*>(x + (int64_t)chunk * xs + uc_flat(g, n, t, h, w) * 8), b);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = a[e] - b[e];
  *reinterpret_cast<uint4*>(d + (int64_t)chunk * ds + uc_flat(gd, n, t, h, w) * 8) = pack8(o);
}
void launch_planar_delta(const __nv_bfloat16* x, int64_t x_plane_stride, UcGeom g, __nv_bfloat16* d, int64_t d_plane_stride, UcGeom gd,
                         int C, cudaStream_t s) {
  const int64_t total = (int64_t)gd.N * gd.T * gd.H * gd.W * (C / 8);
  if (total == 0) return;
  planar_delta_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, x_plane_stride, g, d, d_plane_stride, gd, total);
  count_launch();
}

// MaxPool (1,3,3)/(1,2,2)/pad(0,1,1) in planar layout.  Inputs are post-ReLU (>= 0) and the pads are zero, so reading
// the zero pad is equivalent to the reference's -inf padding except at the far edge, which is bounds-checked.
__global__ void planar_maxpool_kernel(const __nv_bfloat16* __restrict__ x, int64_t xs, UcGeom gi, __nv_bfloat16* __restrict__ y, int64_t ys,
                                      UcGeom go, int64_t total, const __nv_bfloat16* __restrict__ xlo, __nv_bfloat16* __restrict__ ylo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t npos = (int64_t)go.N * go.T * go.H * go.W;
  const int chunk = (int)(i / npos);
  int64_t r = i - (int64_t)chunk * npos;
  const int w = (int)(r % go.W); r /= go.W;
  const int h = (int)(r % go.H); r /= go.H;
  const int t = (int)(r % go.T);
  const int n = (int)(r / go.T);
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
  for (int dh = 0; dh < 3; ++dh) {
    const int hi = 2 * h - 1 + dh;
    if ((unsigned)hi >= (unsigned)gi.H) continue;
    for (int dw = 0; dw < 3; ++dw) {
      const int wi = 2 * w - 1 + dw;
      if ((unsigned)wi >= (unsigned)gi.W) continue;
      float f[8];
      const int64_t src = (int64_t)chunk * xs + uc_flat(gi, n, t, hi, wi) * 8;
      unpack8(*reinterpret_cast<const uint4*>(x + src), f);
      if (xlo) {
        float fl[8];
        unpack8(*reinterpret_cast<const uint4*>(xlo + src), fl);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += fl[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], f[e]);
    }
  }
  const int64_t dst = (int64_t)chunk * ys + uc_flat(go, n, t, h, w) * 8;
  const uint4 hi4 = pack8(m);
  *reinterpret_cast<uint4*>(y + dst) = hi4;
  if (ylo) {
    float hf[8], lo[8];
    unpack8(hi4, hf);
#pragma unroll
    for (int e = 0; e < 8; ++e) lo[e] = m[e] - hf[e];
    *reinterpret_cast<uint4*>(ylo + dst) = pack8(lo);
  }
}
void launch_planar_maxpool(const __nv_bfloat16* x, int64_t x_plane_stride, UcGeom gi, __nv_bfloat16* y, int64_t y_plane_stride, UcGeom go,
                           int C, cudaStream_t s, const __nv_bfloat16* xlo, __nv_bfloat16* ylo) {
  const int64_t total = (int64_t)go.N * go.T * go.H * go.W * (C / 8);
  if (total == 0) return;
  planar_maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, x_plane_stride, gi, y, y_plane_stride, go, total, xlo, ylo);
  count_launch();
}

// One block (32 x 8 threads) per (n, t, band of VR_ROWS image rows): the band (+1 halo row/column on every side) is staged
// in shared memory with coalesced, batched loads and no integer divisions, and each thread
// produces pixel pairs (16-byte units) of both row buffers.
constexpr int VR_ROWS = 8;
// uint8 crops are scaled exactly like the reference: astype(float32) / 255.0 (video.py:552-556)
template <typename T> __device__ __forceinline__ float vr_norm(float x) { return x; }
template <> __device__ __forceinline__ float vr_norm<uint8_t>(float x) { return x / 255.0f; }
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256) video_rows_kernel(const T* __restrict__ video, const float* __restrict__ lapw,
                                                         __nv_bfloat16* __restrict__ xs, __nv_bfloat16* __restrict__ xl, int64_t set_stride,
                                                         UcGeom g, int Tn, int H, int W, int bands, float inv_div) {
  extern __shared__ float tile[];   // [3][VR_ROWS + 2][W + 2]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int WP2 = W + 2, per_c = (VR_ROWS + 2) * WP2;
  const int band = blockIdx.x % bands;
  const int nt = blockIdx.x / bands;            // n * Tn + t
  const int t = nt % Tn, n = nt / Tn;
  const int h0 = band * VR_ROWS;
  __shared__ float lw[81];
  if (ty == 0 && tx < 27) { lw[tx] = lapw[tx]; lw[tx + 27] = lapw[tx + 27]; lw[tx + 54] = lapw[tx + 54]; }
  // tile fill: all loads of a row are issued before the first shared-memory store (the kernel is latency-bound otherwise)
  for (int r = ty; r < VR_ROWS + 2; r += 8) {
    const int hh = h0 + r - 1;
    const bool rin = (unsigned)hh < (unsigned)H;
    if (LAYOUT == 0) {
      const T* rp0 = video + ((((int64_t)n * 3) * Tn + t) * H + (rin ? hh : 0)) * W;
      const int64_t cstride = (int64_t)Tn * H * W;
      for (int col0 = 0; col0 < WP2; col0 += 128) {
        float vals[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ww = col0 + tx + 32 * k - 1;
            vals[c][k] = (rin && (unsigned)ww < (unsigned)W) ? vr_norm<T>((float)rp0[c * cstride + ww]) : 0.f;
          }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int col = col0 + tx + 32 * k;
            if (col < WP2) tile[c * per_c + r * WP2 + col] = vals[c][k];
          }
      }
    } else {
      const T* rp = video + (((int64_t)nt * H + (rin ? hh : 0)) * W) * 3;
      for (int e0 = 0; e0 < 3 * WP2; e0 += 256) {
        float vals[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = e0 + tx + 32 * k, col = e / 3, c = e - col * 3, ww = col - 1;
          vals[k] = (rin && e < 3 * WP2 && (unsigned)ww < (unsigned)W) ? vr_norm<T>((float)rp[ww * 3 + c]) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e = e0 + tx + 32 * k, col = e / 3, c = e - col * 3;
          if (e < 3 * WP2) tile[c * per_c + r * WP2 + col] = vals[k];
        }
      }
    }
  }
  __syncthreads();
  const int h = h0 + ty;
  if (h >= H) return;
  // each thread produces strips of 4 consecutive pixels (two 16-byte units): the 3x6 neighbourhood and every laplacian
  // weight are read from shared memory once per strip instead of once per pixel
  const int64_t row_dst = (int64_t)(h & 1) * set_stride + uc_flat(g, n, t, h >> 1, 0) * 8 + 16;
  for (int p0 = 4 * tx; p0 < W; p0 += 128) {
    float acc[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
    float ctr[4][3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        float x[6];
        const float* trow = &tile[ci * per_c + (ty + kh) * WP2 + p0];
#pragma unroll
        for (int j = 0; j < 6; ++j) x[j] = (p0 + j < WP2) ? trow[j] : 0.f;   // tile columns p0 .. p0+5 <-> pixels p0-1 .. p0+4
        if (kh == 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) ctr[q][ci] = x[q + 1];
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int wi = ((kh * 3 + kw) * 3 + ci) * 3;
          const float w0 = lw[wi], w1 = lw[wi + 1], w2 = lw[wi + 2];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q][0] = fmaf(w0, x[q + kw], acc[q][0]);
            acc[q][1] = fmaf(w1, x[q + kw], acc[q][1]);
            acc[q][2] = fmaf(w2, x[q + kw], acc[q][2]);
          }
        }
      }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p0 + 2 * u >= W) break;
      float px[8], lp[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const bool in = p0 + 2 * u + q < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          px[q * 4 + c] = in ? ctr[2 * u + q][c] : 0.f;
          lp[q * 4 + c] = in ? acc[2 * u + q][c] : 0.f;
        }
        px[q * 4 + 3] = 0.f;
        lp[q * 4 + 3] = 0.f;
      }
      *reinterpret_cast<uint4*>(xs + row_dst + (int64_t)(p0 + 2 * u) * 4) = pack8(px);
      *reinterpret_cast<uint4*>(xl + row_dst + (int64_t)(p0 + 2 * u) * 4) = pack8(lp);
    }
  }
}
void launch_video_rows(const void* video, int dtype, int layout, const float* lapw, __nv_bfloat16* xs, __nv_bfloat16* xl,
                       int64_t set_stride, UcGeom g, int H, int W, cudaStream_t s) {
  const int bands = (H + VR_ROWS - 1) / VR_ROWS;
  const int64_t blocks = (int64_t)g.N * g.T * bands;
  if (blocks == 0) return;
  const size_t smem = (size_t)3 * (VR_ROWS + 2) * (W + 2) * sizeof(float);
#define VR(TT, LL, DIV) video_rows_kernel<TT, LL><<<(unsigned)blocks, dim3(32, 8), smem, s>>>(reinterpret_cast<const TT*>(video), lapw, xs, xl, set_stride, g, g.T, H, W, bands, 1.0f / (DIV))
  if (layout == 0) {
    if (dtype == 0) VR(float, 0, 1.0f); else if (dtype == 1) VR(__half, 0, 1.0f); else if (dtype == 2) VR(__nv_bfloat16, 0, 1.0f); else VR(uint8_t, 0, 255.0f);
  } else {
    if (dtype == 0) VR(float, 1, 1.0f); else if (dtype == 1) VR(__half, 1, 1.0f); else if (dtype == 2) VR(__nv_bfloat16, 1, 1.0f); else VR(uint8_t, 1, 255.0f);
  }
#undef VR
  count_launch();
}

// Generalised deterministic mean (see token_kernels.cuh): fp32 rows and/or planar bf16 rows out.
__global__ void planar_mean2_kernel(const __nv_bfloat16* __restrict__ x, int64_t plane_stride, UcGeom g, float* __restrict__ y32, int ld,
                                    int mode, PlanarOut po, const __nv_bfloat16* __restrict__ xlo) {
  __shared__ float part[256][9];
  const int row = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
  const int count = mode == 1 ? g.T * g.H * g.W : (mode == 0 ? g.H * g.W : g.H);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = tid; i < count; i += blockDim.x) {
    int n, t, h, w;
    if (mode == 2) { n = row / g.W; w = row % g.W; t = 0; h = i; }
    else {
      n = mode == 1 ? row : row / g.T;
      w = i % g.W;
      const int r = i / g.W;
      h = r % g.H;
      t = (mode == 1 ? 0 : row % g.T) + r / g.H;
    }
    float f[8];
    const int64_t src = (int64_t)chunk * plane_stride + uc_flat(g, n, t, h, w) * 8;
    unpack8(*reinterpret_cast<const uint4*>(x + src), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
    if (xlo) {
      unpack8(*reinterpret_cast<const uint4*>(xlo + src), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[tid][e] = acc[e];
  __syncthreads();
  for (int sft = blockDim.x >> 1; sft > 0; sft >>= 1) {
    if (tid < sft) {
#pragma unroll
      for (int e = 0; e < 8; ++e) part[tid][e] += part[tid + sft][e];
    }
    __syncthreads();
  }
  if (tid < 8) {
    const float m = part[0][tid] / (float)count;
    if (y32) y32[(int64_t)row * ld + chunk * 8 + tid] = m;
    if (po.y) {
      const int64_t pos = po.grp > 0 ? ((int64_t)row / po.grp) * po.grp_stride + (row % po.grp) + po.off : (int64_t)row + po.off;
      const __nv_bfloat16 hi = __float2bfloat16_rn(m);
      po.y[(int64_t)chunk * po.plane_stride + pos * 8 + tid] = hi;
      if (po.ylo) po.ylo[(int64_t)chunk * po.plane_stride + pos * 8 + tid] = __float2bfloat16_rn(m - __bfloat162float(hi));
    }
  }
}
void launch_planar_mean2(const __nv_bfloat16* x, int64_t plane_stride, UcGeom g, int C, float* y32, int ld, int mode, PlanarOut po, cudaStream_t s,
                         const __nv_bfloat16* xlo) {
  const int rows = mode == 1 ? g.N : (mode == 0 ? g.N * g.T : g.N * g.W);
  if (rows == 0) return;
  const int count = mode == 1 ? g.T * g.H * g.W : (mode == 0 ? g.H * g.W : g.H);
  int threads = 32;
  while (threads < 256 && threads < count) threads <<= 1;
  planar_mean2_kernel<<<dim3(rows, C / 8), threads, 0, s>>>(x, plane_stride, g, y32, ld, mode, po, xlo);
  count_launch();
}

}  // namespace lsd
