// Fused temporal transformer for sm_100a: the four pre-norm encoder layers of TemporalTransformer.forward
// (app/models/temporal.py:64-77: LN -> MHA(8 heads, hd 32) -> +res -> LN -> Linear 256->1024 -> GELU -> Linear 1024->256 -> +res)
// in one launch.  Replaces 28 launches per forward (4 x [LN, in-proj GEMM, attention core, out-proj GEMM, LN, FF1 GEMM, FF2 GEMM]).
//
// One CTA owns G windows (rows = tokens; every window gets a slot of SL = 32 or 64 rows of the 128-row tile, so a warp's 32
// rows never straddle windows and a window's keys start on a K16 boundary: a window's result does not depend on its slot or
// on its co-tenants, bit for bit).  Everything a layer touches stays on the SM:
//   * the residual stream X (128 x 256 fp32) lives in TMEM columns 0..255 for the whole kernel; the out-projection and the
//     second FFN GEMM accumulate straight into it (tcgen05.mma with accumulate), their biases are carried as a cumulative
//     vector added wherever X is read (LayerNorm, final store);
//   * TMEM columns 256..511 hold the transient accumulators: Q|K|V of a head pair (192 columns), the two score tiles
//     S = Q K^T (2 x 112), the two P V products (2 x 32), an FF1 chunk (128);
//   * GEMM operands are fp16 in shared memory, K-major core-matrix layout (planes of 8 channels, rows 16 B apart, no swizzle):
//     LayerNorm output (K = 256), Q / K / V^T of the head pair, the un-normalised probabilities P (flash-style: the row sum is
//     kept in a register and applied to P V), the attention output of the head pair (K = 64), a GELU'd FF chunk (K = 128);
//   * weights (fp16, 1.5 MB per layer) stream from L2 through a 3-slot ring of bulk async copies (TMA engine) in exactly the
//     order the MMA warp consumes them (tok_fused.cuh: tf_layer_blocks).
// fp16 operands with fp32 accumulation: one MMA pass instead of the three of the split-bf16 GEMMs this replaces (logit
// deviation from the fp32 reference arithmetic measured on the CPU: 2.8e-3, tests hold the bf16-route budget of 2e-2).
//
// Warp roles (320 threads): warps 0-7 compute (warp w: TMEM lane quarter w % 4, thread = row; the two warps of a quarter split
// columns, or take one head each in the attention phases), warp 8 streams weights, warp 9 issues every tcgen05.mma.
// Compute and MMA phases alternate strictly (two mbarriers): simple to reason about, and the only latency it exposes is the
// hand-over (~1 us per phase pair; 22 pairs per layer).
#include "tok_fused.cuh"

#include <cstdio>

#include "lsd_kernels.h"
#include "tok_common.cuh"
#include "umma.cuh"

namespace lsd {

using namespace umma;
using namespace tokc;

namespace {

constexpr int TF_CWARPS = 16;                        // compute warps: 4 per TMEM lane quarter
constexpr int TF_THREADS = (TF_CWARPS + 2) * 32;     // + weight producer + MMA issuer
constexpr int TF_RING = 5;
constexpr uint32_t OFF_ALN = 0;                      // LayerNorm output, K = 256: 32 planes
constexpr uint32_t OFF_QK = 65536;                   // Q0 K0 Q1 K1: 4 planes each (hd = 32); aliased by ATT (K = 64: 8 planes) after S
constexpr uint32_t OFF_VT = OFF_QK + 32768;          // V^T of the two heads: [key plane (16)][32 hd rows][8 keys], 8 KB per head
constexpr uint32_t OFF_P = OFF_VT + 16384;           // P of the two heads, keys relative to the row's own window: <= 8 planes (16 KB) each
constexpr uint32_t P_HEAD = 8 * PLANE;
constexpr uint32_t OFF_FF = OFF_VT;                  // GELU'd FF chunk (K = 128: 16 planes = 32 KB) aliases V^T + the first P buffer
constexpr uint32_t OFF_RED = OFF_P;                  // LayerNorm partial sums (4 KB) alias P
constexpr uint32_t OFF_RING = OFF_P + 2 * P_HEAD;
constexpr uint32_t TF_SMEM = OFF_RING + TF_RING * TF_SLOT_BYTES;
static_assert(OFF_RING - OFF_FF >= 16 * PLANE, "the FF chunk operand must fit in the V^T + P region");
static_assert(TF_SMEM <= 225 * 1024, "shared-memory budget");
constexpr uint32_t X_COL = 0, ACC_COL = 256;

}  // namespace

__global__ void __launch_bounds__(TF_THREADS, 1) tok_fused_kernel(const __grid_constant__ TokFusedP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t w_full[TF_RING], w_empty[TF_RING], bar_mma, bar_cmp, bar_f1[2], bar_gd[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  constexpr int W_PROD = TF_CWARPS, W_MMA = TF_CWARPS + 1;

  if (tid == 0) {
    for (int i = 0; i < TF_RING; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_cmp, TF_CWARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_f1[i], 1); mbar_init(&bar_gd[i], TF_CWARPS); }
    fence_barrier_init();
  }
  if (warp == W_MMA) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int G = p.G, SL = p.SL, KW = p.KW;

  if (warp == W_PROD) {
    // ------------------------------------------------------------------ weight producer (one lane)
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w);
      int s = 0;
      for (int l = 0; l < TF_LAYERS; ++l) {
        const uint8_t* ls = src + (size_t)l * p.layer_bytes;
        uint32_t off = 0;
        for (int i = 0; i < p.n_stage_layer; ++i, ++s) {
          const int slot = s % TF_RING;
          const uint32_t ph = (uint32_t)(s / TF_RING) & 1u;
          const uint32_t bytes = __ldg(p.stage_bytes + i);
          mbar_wait(&w_empty[slot], ph ^ 1u);
          mbar_arrive_expect_tx(&w_full[slot], bytes);
          bulk_s2(sbase + OFF_RING + (uint32_t)slot * TF_SLOT_BYTES, ls + off, bytes, &w_full[slot]);
          off += bytes;
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer (whole warp runs the loop, one lane issues)
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint64_t desc_hi = ((uint64_t)8 << 32) | ((uint64_t)1 << 46);   // SBO = 128 B, descriptor version 1
    uint32_t par_c = 0, par_g[2] = {0u, 0u};
    int ws = 0;                                                           // weight stages consumed so far
    // debug (p.dbg != nullptr, CTA 0): clock64 at the start and at the issue-end of every MMA phase
    long long* dbg = (p.dbg && blockIdx.x == 0 && lane == 0) ? reinterpret_cast<long long*>(p.dbg) : nullptr;
    int dbg_n = 0;
    auto wait_cmp = [&]() { mbar_wait(&bar_cmp, par_c); par_c ^= 1u; tc_fence_after(); if (dbg && dbg_n < 96) dbg[2 * dbg_n] = clock64(); };
    auto done = [&]() { mma_commit_pred(&bar_mma, leader); if (dbg && dbg_n < 96) { dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; } };
    auto desc = [&](uint32_t byte_addr, uint32_t lbo_bytes) -> uint64_t {
      return desc_hi | (uint64_t)(((lbo_bytes >> 4) & 0x3FFFu) << 16) | (uint64_t)((byte_addr >> 4) & 0x3FFFu);
    };
    // D[tmem_d] (+)= A[a_addr: planes of 128 rows] * W^T, W = (N x 16*k16) streamed through the ring in stages of kps K16 steps
    auto gemm_w = [&](uint32_t a_addr, int N, int k16, int kps, uint32_t tmem_d, uint32_t acc) {
      const uint32_t idesc = idesc_f16(128, N);
      for (int k0 = 0; k0 < k16; k0 += kps, ++ws) {
        const int slot = ws % TF_RING;
        mbar_wait(&w_full[slot], (uint32_t)(ws / TF_RING) & 1u);
        tc_fence_after();
        const uint32_t wb = sbase + OFF_RING + (uint32_t)slot * TF_SLOT_BYTES;
#pragma unroll 1
        for (int j = 0; j < kps; ++j) {
          const uint64_t da = desc(a_addr + (uint32_t)(k0 + j) * 2u * PLANE, PLANE);
          const uint64_t db = desc(wb + (uint32_t)j * (uint32_t)N * 32u, (uint32_t)N * 16u);
          mma_bf16_ss_pred(tmem_d, da, db, idesc, acc, leader);
          acc = 1u;
        }
        mma_commit_pred(&w_empty[slot], leader);
      }
    };
    const uint32_t X = tmem + X_COL, ACC = tmem + ACC_COL;
    // S_{h,slot} = Q_h K_{h,slot}^T: all 128 query rows against the KW keys of window `slot` (rows of other windows produce
    // values nobody reads); hd = 32: two K16 steps.  Score tile of (h, slot) at ACC + h*128 + slot*SL.
    auto mma_scores = [&]() {
      const uint32_t idesc = idesc_f16(128, KW);
#pragma unroll 1
      for (int h = 0; h < 2; ++h)
#pragma unroll 1
        for (int sl = 0; sl < G; ++sl)
#pragma unroll 1
          for (int j = 0; j < 2; ++j) {
            const uint64_t da = desc(sbase + OFF_QK + (uint32_t)(2 * h) * 8192u + (uint32_t)j * 2u * PLANE, PLANE);
            const uint64_t db = desc(sbase + OFF_QK + (uint32_t)(2 * h + 1) * 8192u + (uint32_t)j * 2u * PLANE + (uint32_t)(sl * SL) * 16u, PLANE);
            mma_bf16_ss_pred(ACC + (uint32_t)(h * 128 + sl * SL), da, db, idesc, j ? 1u : 0u, leader);
          }
    };
    // O_{h,slot} = P_h V_{h,slot}: P holds, for every row, the probabilities over the keys of the row's OWN window (K = KW), so
    // the product with window `slot`'s V is meaningful for that window's rows only; O of (h, slot) at ACC + (h*G + slot)*32.
    auto mma_pv = [&]() {
      const uint32_t idesc = idesc_f16(128, 32);
#pragma unroll 1
      for (int h = 0; h < 2; ++h)
#pragma unroll 1
        for (int sl = 0; sl < G; ++sl)
#pragma unroll 1
          for (int j = 0; j < KW / 16; ++j) {
            const uint64_t da = desc(sbase + OFF_P + (uint32_t)h * P_HEAD + (uint32_t)j * 2u * PLANE, PLANE);
            const uint64_t db = desc(sbase + OFF_VT + (uint32_t)h * 8192u + (uint32_t)(sl * SL / 8 + 2 * j) * 512u, 512u);
            mma_bf16_ss_pred(ACC + (uint32_t)((h * G + sl) * 32), da, db, idesc, j ? 1u : 0u, leader);
          }
    };
    for (int l = 0; l < TF_LAYERS; ++l) {
      for (int hp = 0; hp < 4; ++hp) {
        wait_cmp();                                                       // LN1 output (hp = 0) / attention output of pair hp-1
        if (hp > 0) gemm_w(sbase + OFF_QK, 256, 4, 2, X, 1u);             // X += att_{hp-1} @ Wo[:, 64(hp-1) : 64hp]^T
        gemm_w(sbase + OFF_ALN, 192, 16, 2, ACC, 0u);                     // Q|K|V of heads 2hp, 2hp+1
        done();
        wait_cmp();                                                       // Q, K, V^T in shared memory
        mma_scores();
        done();
        wait_cmp();                                                       // P in shared memory
        mma_pv();
        done();
      }
      wait_cmp();                                                         // attention output of the last pair
      gemm_w(sbase + OFF_QK, 256, 4, 2, X, 1u);
      done();
      // ---- FFN, software-pipelined over chunks of 128 hidden channels: FF1 of chunk c+2 and FF2 of chunk c are issued while the
      // compute warps apply the GELU to chunk c+1 (two accumulator halves, two GELU'd operands; the barriers of the two halves
      // alternate, so neither side can run more than one completed phase ahead of the other)
      wait_cmp();                                                         // LN2 output
      gemm_w(sbase + OFF_ALN, 128, 16, 4, ACC, 0u);                       // h_0 = LN2 @ W1[0:128]^T
      mma_commit_pred(&bar_f1[0], leader);
      gemm_w(sbase + OFF_ALN, 128, 16, 4, ACC + 128u, 0u);                // h_1
      mma_commit_pred(&bar_f1[1], leader);
      for (int c = 0; c < 8; ++c) {
        const int hb = c & 1;
        mbar_wait(&bar_gd[hb], par_g[hb]); par_g[hb] ^= 1u; tc_fence_after();            // GELU'd chunk c in operand hb, accumulator half hb read
        gemm_w(sbase + (hb ? OFF_FF : OFF_QK), 256, 8, 2, X, 1u);         // X += h_c @ W2[:, 128c : 128c+128]^T
        if (c + 2 < 8) {
          gemm_w(sbase + OFF_ALN, 128, 16, 4, ACC + (uint32_t)hb * 128u, 0u);   // h_{c+2}
          mma_commit_pred(&bar_f1[hb], leader);
        }
      }
      done();
    }
  } else {
    // ------------------------------------------------------------------ compute warps: lane quarter q, column quarter cq
    // Tile row (= TMEM lane) -> token: a window owns 4/G lane quarters and its tokens are interleaved over them (token = lane *
    // (4/G) + quarter in window).  A warp can only read the TMEM lanes of quarter (warp % 4), and warp w issues on scheduler w % 4,
    // so with a window's tokens in consecutive rows the quarters that hold a slot's unused rows (31 of 64 at NT = 33) would leave
    // their schedulers idle while the others do all the elementwise work; interleaved, every scheduler gets the same share, and a
    // warp's rows still belong to ONE window (the TMEM address of a tcgen05.ld must be warp-uniform).  Operand rows that are
    // indexed by KEY (K, V^T) are stored in canonical order (krow): the score / P V tiles of a window then are contiguous column
    // ranges, whatever the order of the query rows.
    const int q = warp & 3, cq = warp >> 2;
    const int row = q * 32 + lane;
    const int QW = 4 / G;
    const int slot = q / QW;
    const int lrow = lane * QW + (q - slot * QW);
    const int krow = slot * SL + lrow;
    const int win = blockIdx.x * G + slot;
    const bool valid = slot < G && win < p.B && lrow < p.NT;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t X = lane_base + X_COL, ACC = lane_base + ACC_COL;
    const uint32_t row16 = (uint32_t)row * 16u;
    float* grow = p.tok + ((size_t)(valid ? win : 0) * p.NT + (valid ? lrow : 0)) * TF_D;
    uint32_t par_m = 0, par_f[2] = {0u, 0u};
    // debug (p.dbg != nullptr, CTA 0, warp 0): clock64 when an accumulator arrives and when the compute phase hands over
    long long* dbg = (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0) ? reinterpret_cast<long long*>(p.dbg) + 192 : nullptr;
    int dbg_n = 0;
    if (dbg) dbg[0] = clock64();
    auto wait_mma = [&]() { mbar_wait(&bar_mma, par_m); par_m ^= 1u; tc_fence_after(); if (dbg && dbg_n < 96) dbg[2 * dbg_n] = clock64(); };
    auto done = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_cmp);
      if (dbg && dbg_n < 96) { dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; }
    };
    auto quarter_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory"); };   // the four warps of this lane quarter
    const float* cum_all = p.vec + (size_t)TF_LAYERS * TF_VEC_LAYER;

    // X <- tok rows (this warp's 64 columns); rows outside the batch are zero
#pragma unroll 1
    for (int c0 = cq * 64; c0 < cq * 64 + 64; c0 += 32) {
      float v[32];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 t = valid ? *reinterpret_cast<const float4*>(grow + c0 + 4 * e) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * e] = t.x; v[4 * e + 1] = t.y; v[4 * e + 2] = t.z; v[4 * e + 3] = t.w;
      }
      tmem_st32(X + (uint32_t)c0, v);
    }
    tmem_st_wait();

    // LayerNorm of (X + cum) -> A_ln (fp16 planes).  A warp holds its 64 columns of the row in registers; the row statistics
    // are combined across the four warps of the lane quarter through shared memory in a fixed order (two-pass variance).
    auto layer_norm = [&](const float* cum, const float* g, const float* b) {
      const int c0 = cq * 64;
      const uint32_t red = sbase + OFF_RED + (uint32_t)row * 4u;           // [2][4 cq][128 rows] floats
      float s = 0.f;
#pragma unroll 1
      for (int c = c0; c < c0 + 64; c += 32) {
        float v[32];
        tmem_ld32(X + (uint32_t)c, v);
        tmem_ld_wait();
        add32(v, cum + c);
#pragma unroll
        for (int e = 0; e < 32; ++e) s += v[e];
      }
      st_shared_f32(red + (uint32_t)cq * 512u, s);
      quarter_sync();
      const float mu = ((ld_shared_f32(red) + ld_shared_f32(red + 512u)) + (ld_shared_f32(red + 1024u) + ld_shared_f32(red + 1536u))) * (1.0f / TF_D);
      float d2 = 0.f;
#pragma unroll 1
      for (int c = c0; c < c0 + 64; c += 32) {
        float v[32];
        tmem_ld32(X + (uint32_t)c, v);
        tmem_ld_wait();
        add32(v, cum + c);
#pragma unroll
        for (int e = 0; e < 32; ++e) { const float d = v[e] - mu; d2 = fmaf(d, d, d2); }
      }
      st_shared_f32(red + 2048u + (uint32_t)cq * 512u, d2);
      quarter_sync();
      const float var = ((ld_shared_f32(red + 2048u) + ld_shared_f32(red + 2560u)) + (ld_shared_f32(red + 3072u) + ld_shared_f32(red + 3584u))) * (1.0f / TF_D);
      const float rstd = valid ? 1.0f / sqrtf(var + 1e-5f) : 0.f;          // rows outside the batch: zeros
#pragma unroll 1
      for (int c = c0; c < c0 + 64; c += 32) {
        float v[32];
        tmem_ld32(X + (uint32_t)c, v);
        tmem_ld_wait();
        add32(v, cum + c);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c) + e4), bb = __ldg(reinterpret_cast<const float4*>(b + c) + e4);
          v[4 * e4] = valid ? fmaf((v[4 * e4] - mu) * rstd, gg.x, bb.x) : 0.f;
          v[4 * e4 + 1] = valid ? fmaf((v[4 * e4 + 1] - mu) * rstd, gg.y, bb.y) : 0.f;
          v[4 * e4 + 2] = valid ? fmaf((v[4 * e4 + 2] - mu) * rstd, gg.z, bb.z) : 0.f;
          v[4 * e4 + 3] = valid ? fmaf((v[4 * e4 + 3] - mu) * rstd, gg.w, bb.w) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) st_plane8(sbase + OFF_ALN + (uint32_t)(c / 8 + j) * PLANE + row16, v + 8 * j);
      }
    };

    for (int l = 0; l < TF_LAYERS; ++l) {
      const float* lv = p.vec + (size_t)l * TF_VEC_LAYER;
      if (l > 0) wait_mma();                                              // FF2 of the previous layer complete
      layer_norm(cum_all + (size_t)(2 * l) * TF_D, lv, lv + 256);
      done();
      const float* qkv_b = lv + 1024;
      float inv_sum = 0.f;
      for (int hp = 0; hp < 4; ++hp) {
        // ---- Q|K|V epilogue: six 32-column blocks [Q0 K0 V0 Q1 K1 V1] of the accumulator; +bias, Q scaled by 1/sqrt(32), fp16
        // planes; V transposed (B operand of P V).  cq 0: Q0 K0, cq 1: Q1 K1, cq 2: V0, cq 3: V1 (the transposing blocks alone).
        wait_mma();
        {
          const int b0 = cq == 0 ? 0 : (cq == 1 ? 3 : (cq == 2 ? 2 : 5));
          const int nb = cq < 2 ? 2 : 1;
#pragma unroll 1
          for (int bi = 0; bi < nb; ++bi) {
            const int blk = b0 + bi, head = blk / 3, part = blk - head * 3;
            float v[32];
            tmem_ld32(ACC + (uint32_t)(blk * 32), v);
            tmem_ld_wait();
            add32(v, qkv_b + hp * 192 + blk * 32);
            const float sc = !valid ? 0.f : (part == 0 ? 0.17677669529663688f : 1.0f);
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] *= sc;
            if (part < 2) {
              const uint32_t base = sbase + OFF_QK + (uint32_t)(2 * head + part) * 8192u + (part == 0 ? row16 : (uint32_t)krow * 16u);
#pragma unroll
              for (int j = 0; j < 4; ++j) st_plane8(base + (uint32_t)j * PLANE, v + 8 * j);
            } else {
              const uint32_t base = sbase + OFF_VT + (uint32_t)head * 8192u + (uint32_t)(krow >> 3) * 512u + (uint32_t)(krow & 7) * 2u;
#pragma unroll
              for (int d = 0; d < 32; ++d) st_shared_u16(base + (uint32_t)d * 16u, __half_as_ushort(__float2half_rn(v[d])));
            }
          }
        }
        done();
        // ---- softmax over the row's own window (score columns slot*SL .. +NT of head cq), un-normalised P -> fp16 planes,
        // keys relative to the window; one TMEM pass (NT <= 64 values per row in registers)
        wait_mma();
        if (cq < 2) {
          const uint32_t S = ACC + (uint32_t)(cq * 128 + slot * SL);
          float mx = -INFINITY;
#pragma unroll 1
          for (int c0 = 0; c0 < p.NT; c0 += 32) {
            float v[32];
            tmem_ld32(S + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) if (c0 + e < p.NT) mx = fmaxf(mx, v[e]);
          }
          float sum = 0.f;
          const uint32_t pbase = sbase + OFF_P + (uint32_t)cq * P_HEAD + row16;
#pragma unroll 1
          for (int c0 = 0; c0 < KW; c0 += 32) {
            float v[32];
            tmem_ld32(S + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float pe = (valid && c0 + e < p.NT) ? __expf(v[e] - mx) : 0.f;
              // the row sum is taken over the fp16-rounded probabilities the tensor core multiplies with V
              const float pr = __half2float(__float2half_rn(pe));
              sum += pr;
              v[e] = pr;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c0 + 8 * j < KW) st_plane8(pbase + (uint32_t)(c0 / 8 + j) * PLANE, v + 8 * j);
          }
          inv_sum = (valid && sum > 0.f) ? 1.0f / sum : 0.f;
        }
        done();
        // ---- attention output of head 2hp + cq: (P V) / rowsum -> planes 4cq..4cq+3 of the K = 64 operand (aliases Q/K)
        wait_mma();
        if (cq < 2) {
          float v[32];
          tmem_ld32(ACC + (uint32_t)((cq * G + slot) * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] *= inv_sum;
#pragma unroll
          for (int j = 0; j < 4; ++j) st_plane8(sbase + OFF_QK + (uint32_t)(4 * cq + j) * PLANE + row16, v + 8 * j);
        }
        done();
      }
      // ---- LN2 (X now includes the attention update; its bias rides in the cumulative vector)
      wait_mma();
      layer_norm(cum_all + (size_t)(2 * l + 1) * TF_D, lv + 512, lv + 768);
      done();
      // ---- FFN chunks: GELU(acc + b1) -> fp16 planes of the K = 128 operand (this warp: 32 of the 128 columns); accumulator half
      // and operand buffer alternate with the chunk (see the MMA warp)
      const float* b1 = lv + 1024 + 768;
      for (int c = 0; c < 8; ++c) {
        const int hb = c & 1;
        mbar_wait(&bar_f1[hb], par_f[hb]); par_f[hb] ^= 1u; tc_fence_after();
        {
          float v[32];
          tmem_ld32(ACC + (uint32_t)(hb * 128 + cq * 32), v);
          tmem_ld_wait();
          add32(v, b1 + c * 128 + cq * 32);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = valid ? gelu_fast(v[e]) : 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) st_plane8(sbase + (hb ? OFF_FF : OFF_QK) + (uint32_t)(cq * 4 + j) * PLANE + row16, v + 8 * j);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_gd[hb]);
      }
    }
    // ---- tok <- X + total bias
    wait_mma();
    if (__any_sync(0xffffffffu, valid)) {     // warp-uniform: tcgen05.ld is .sync.aligned, only the stores are per row
      const float* cum = cum_all + (size_t)(2 * TF_LAYERS) * TF_D;
#pragma unroll 1
      for (int c0 = cq * 64; c0 < cq * 64 + 64; c0 += 32) {
        float v[32];
        tmem_ld32(X + (uint32_t)c0, v);
        tmem_ld_wait();
        add32(v, cum + c0);
        if (valid) {
#pragma unroll
          for (int e = 0; e < 8; ++e) *reinterpret_cast<float4*>(grow + c0 + 4 * e) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

cudaError_t tok_fused_device_init() {
  return cudaFuncSetAttribute(tok_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM);
}

// Row slot per window: a multiple of 32 (a warp's rows belong to one window).  KW = key extent per window (multiple of 16).
void tok_fused_geometry(int NT, int& SL, int& G, int& KW) {
  SL = NT <= 32 ? 32 : 64;
  G = 128 / SL;
  KW = (NT + 15) / 16 * 16;
}
bool tok_fused_supported(int NT) { return NT >= 1 && NT <= 64; }

void launch_tok_fused(const TokFusedP& p, cudaStream_t s) {
  const int grid = (p.B + p.G - 1) / p.G;
  if (grid <= 0) return;
  tok_fused_kernel<<<grid, TF_THREADS, TF_SMEM, s>>>(p);
  count_launch();
}

}  // namespace lsd
