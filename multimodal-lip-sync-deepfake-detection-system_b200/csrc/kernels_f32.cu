// fp32 CUDA-core kernels of the window-scoring path (parity path, <= 1e-4 relative on logits) plus the
// memory-bound glue kernels that both precisions share.  All activations are channels-last.
// Reference semantics: app/models/*.py of the reference (cited per kernel); SURVEY.md App. A.
#include "lsd_kernels.h"

#include <atomic>
#include <math.h>

namespace lsd {

static std::atomic<int64_t> g_launches{0};
int64_t kernel_launches() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_GELU) return gelu_erf(v);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution, M = N*To*Ho*Wo output positions, N = Cout, K = taps*Cin.
// 128x64 tile, BK = 16, 256 threads, 8x4 register tile per thread.  Fixed reduction order per output
// (tap-major, channel-minor) so results do not depend on batch composition.
// Replaces nn.Conv3d/Conv2d/Conv1d/Linear + eval BatchNorm + activation + residual of the reference.
template <bool VEC>
__global__ void __launch_bounds__(256) conv_f32_kernel(const ConvF32 p) {
  constexpr int BM = 128, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int rn[BM], rt[BM], rh[BM], rw[BM];
  const int tid = threadIdx.x;
  const int64_t M = (int64_t)p.N * p.To * p.Ho * p.Wo;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  if (tid < BM) {
    int64_t m = m0 + tid;
    if (m < M) {
      int wo = (int)(m % p.Wo);
      int64_t r = m / p.Wo;
      int ho = (int)(r % p.Ho);
      r /= p.Ho;
      int to = (int)(r % p.To);
      rn[tid] = (int)(r / p.To);
      rt[tid] = to * p.st - p.pt;
      rh[tid] = ho * p.sh - p.ph;
      rw[tid] = wo * p.sw - p.pw;
    } else {
      rn[tid] = -1; rt[tid] = 0; rh[tid] = 0; rw[tid] = 0;
    }
  }
  __syncthreads();
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const int tm = tid >> 4, tn = tid & 15;
  const int Ktot = p.kt * p.kh * p.kw * p.Cin;
  const bool wvec = (p.w_ld & 3) == 0;
  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    if (VEC) {
      const int tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
      const int dkw = tap % p.kw, dkh = (tap / p.kw) % p.kh, dkt = tap / (p.kw * p.kh);
      const int kq = tid & 3;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = (tid >> 2) + 64 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n = rn[row];
        const int ti = rt[row] + dkt, hi = rh[row] + dkh, wi = rw[row] + dkw;
        if (n >= 0 && (unsigned)ti < (unsigned)p.Ti && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi) {
          const int64_t pix = (((int64_t)n * p.Ti + ti) * p.Hi + hi) * p.Wi + wi;
          v = *reinterpret_cast<const float4*>(p.x + pix * p.in_ld + c0 + kq * 4);
        }
        As[kq * 4 + 0][row] = v.x; As[kq * 4 + 1][row] = v.y; As[kq * 4 + 2][row] = v.z; As[kq * 4 + 3][row] = v.w;
      }
    } else {
      const int kk = tid & 15, k = k0 + kk;
      const bool kval = k < Ktot;
      const int tap = kval ? k / p.Cin : 0, c = k - tap * p.Cin;
      const int dkw = tap % p.kw, dkh = (tap / p.kw) % p.kh, dkt = tap / (p.kw * p.kh);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (tid >> 4) + 16 * i;
        float v = 0.f;
        const int n = rn[row];
        const int ti = rt[row] + dkt, hi = rh[row] + dkh, wi = rw[row] + dkw;
        if (kval && n >= 0 && (unsigned)ti < (unsigned)p.Ti && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi) {
          const int64_t pix = (((int64_t)n * p.Ti + ti) * p.Hi + hi) * p.Wi + wi;
          v = p.x[pix * p.in_ld + c];
        }
        As[kk][row] = v;
      }
    }
    {
      const int kk = tid >> 4, col = (tid & 15) * 4, k = k0 + kk;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < Ktot) {
        const float* wp = p.w + (int64_t)k * p.w_ld + n0 + col;
        if (wvec && n0 + col + 3 < p.Cout) {
          v = *reinterpret_cast<const float4*>(wp);
        } else {
          if (n0 + col + 0 < p.Cout) v.x = wp[0];
          if (n0 + col + 1 < p.Cout) v.y = wp[1];
          if (n0 + col + 2 < p.Cout) v.z = wp[2];
          if (n0 + col + 3 < p.Cout) v.w = wp[3];
        }
      }
      *reinterpret_cast<float4*>(&Bs[kk][col]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tm * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue: y = act(acc*scale + shift + res)
  float sc[4], sf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = n0 + tn * 4 + j;
    sc[j] = (p.scale && c < p.Cout) ? p.scale[c] : 1.0f;
    sf[j] = (p.shift && c < p.Cout) ? p.shift[c] : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + tm * 8 + i;
    if (m >= M) continue;
    int64_t row = m;
    if (p.grp > 0) row = (m / p.grp) * p.grp_stride + (m % p.grp) + p.row_off;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tn * 4 + j;
      if (c >= p.Cout) continue;
      float v = fmaf(acc[i][j], sc[j], sf[j]);
      if (p.res) v += p.res[m * p.res_ld + c];
      p.y[row * p.out_ld + c] = apply_act(v, p.act);
    }
  }
}

void launch_conv_f32(const ConvF32& p, cudaStream_t s) {
  const int64_t M = (int64_t)p.N * p.To * p.Ho * p.Wo;
  if (M <= 0) return;
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((p.Cout + 63) / 64));
  const bool vec = (p.Cin % 16 == 0) && (p.in_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
  if (vec) conv_f32_kernel<true><<<grid, 256, 0, s>>>(p);
  else conv_f32_kernel<false><<<grid, 256, 0, s>>>(p);
  count_launch();
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_as_f32(const void* src, int dtype, int64_t i) {
  switch (dtype) {
    case 0: return reinterpret_cast<const float*>(src)[i];
    case 1: return __half2float(reinterpret_cast<const __half*>(src)[i]);
    case 2: return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
    default: return (float)reinterpret_cast<const uint8_t*>(src)[i];
  }
}

// (N,C,T,H,W) -> (N,T,H,W,C): one thread per output pixel; reads are coalesced per channel plane.
__global__ void video_to_ndhwc_kernel(const void* src, int dtype, float* dst, int64_t npix_per_n, int C, int64_t total_pix) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pix) return;
  const int64_t n = i / npix_per_n, r = i - n * npix_per_n;
  const float div = dtype == 3 ? 255.0f : 1.0f;  // uint8 crops: astype(float32) / 255.0 (video.py:552-556)
  for (int c = 0; c < C; ++c) dst[i * C + c] = load_as_f32(src, dtype, (n * C + c) * npix_per_n + r) / div;
}
void launch_video_to_ndhwc(const void* src, int dtype, float* dst, int N, int C, int T, int H, int W, cudaStream_t s) {
  const int64_t per = (int64_t)T * H * W, total = per * N;
  if (total == 0) return;
  video_to_ndhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, dtype, dst, per, C, total);
  count_launch();
}

__global__ void cast_to_f32_kernel(const void* src, int dtype, float* dst, int64_t n, float div) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = load_as_f32(src, dtype, i) / div;
}
void launch_cast_to_f32(const void* src, int dtype, float* dst, int64_t n, float div, cudaStream_t s) {
  if (n == 0) return;
  cast_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dtype, dst, n, div);
  count_launch();
}

// window w, frame t <- track[starts[w] + t] / 255 (video.py:552-556: astype(float32) / 255.0)
__global__ void gather_windows_u8_kernel(const uint8_t* track, int n_frames, const int32_t* starts, float* dst,
                                         int T, int frame_elems, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = (int)(i % frame_elems);
  const int64_t r = i / frame_elems;
  const int t = (int)(r % T);
  const int w = (int)(r / T);
  int f = starts[w] + t;
  f = f < 0 ? 0 : (f >= n_frames ? n_frames - 1 : f);
  dst[i] = (float)track[(int64_t)f * frame_elems + e] / 255.0f;
}
void launch_gather_windows_u8(const uint8_t* track, int n_frames, const int32_t* starts, float* dst, int n, int T,
                              int frame_elems, cudaStream_t s) {
  const int64_t total = (int64_t)n * T * frame_elems;
  if (total == 0) return;
  gather_windows_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(track, n_frames, starts, dst, T, frame_elems, total);
  count_launch();
}

// predictor.py:525-552: columns a_start..a_start+Ta, repeating the last available column past the end.
__global__ void gather_audio_kernel(const float* mel, int F, int Ta_full, const int32_t* a_starts, float* dst, int Ta, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % Ta);
  const int64_t r = i / Ta;
  const int f = (int)(r % F);
  const int w = (int)(r / F);
  int col = a_starts[w] + c;
  col = col >= Ta_full ? Ta_full - 1 : col;
  col = col < 0 ? 0 : col;
  dst[i] = mel[(int64_t)f * Ta_full + col];
}
void launch_gather_audio(const float* mel_full, int F, int Ta_full, const int32_t* a_starts, float* dst, int n, int Ta, cudaStream_t s) {
  const int64_t total = (int64_t)n * F * Ta;
  if (total == 0) return;
  gather_audio_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(mel_full, F, Ta_full, a_starts, dst, Ta, total);
  count_launch();
}

// MaxPool (1,3,3)/(1,2,2)/pad(0,1,1) (visual_encoder.py:124-128) and MaxPool2d 3/2/1 (audio_encoder.py:139); -inf padding.
__global__ void maxpool3x3s2_kernel(const float* x, float* y, int Hi, int Wi, int Ho, int Wo, int C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t r = i / C;
  const int wo = (int)(r % Wo); r /= Wo;
  const int ho = (int)(r % Ho);
  const int64_t f = r / Ho;
  float m = -INFINITY;
  for (int dh = 0; dh < 3; ++dh) {
    const int hi = ho * 2 - 1 + dh;
    if ((unsigned)hi >= (unsigned)Hi) continue;
    for (int dw = 0; dw < 3; ++dw) {
      const int wi = wo * 2 - 1 + dw;
      if ((unsigned)wi >= (unsigned)Wi) continue;
      m = fmaxf(m, x[((f * Hi + hi) * Wi + wi) * C + c]);
    }
  }
  y[i] = m;
}
void launch_maxpool3x3s2(const float* x, float* y, int frames, int Hi, int Wi, int C, cudaStream_t s) {
  const int Ho = (Hi + 2 - 3) / 2 + 1, Wo = (Wi + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)frames * Ho * Wo * C;
  if (total == 0) return;
  maxpool3x3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, Hi, Wi, Ho, Wo, C, total);
  count_launch();
}

// mean over R of x[a][r][b]: block = 32 channels x 8 row-lanes, fixed-order tree => deterministic.
__global__ void mean_mid_kernel(const float* x, float* y, int R, int Bc, int out_ld) {
  __shared__ float part[8][33];
  const int a = blockIdx.x;
  const int b = blockIdx.y * 32 + threadIdx.x;
  float acc = 0.f;
  if (b < Bc) {
    const float* base = x + (int64_t)a * R * Bc + b;
    for (int r = threadIdx.y; r < R; r += 8) acc += base[(int64_t)r * Bc];
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && b < Bc) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    y[(int64_t)a * out_ld + b] = t / (float)R;
  }
}
void launch_mean_mid(const float* x, float* y, int A, int R, int Bc, int out_ld, cudaStream_t s) {
  if (A == 0 || Bc == 0) return;
  dim3 grid(A, (Bc + 31) / 32), block(32, 8);
  mean_mid_kernel<<<grid, block, 0, s>>>(x, y, R, Bc, out_ld);
  count_launch();
}

// artifact_detector.py:167: v_map[:,:,1:] - v_map[:,:,:-1]
__global__ void delta_t_kernel(const float* x, float* y, int T, int64_t S, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t per = (int64_t)(T - 1) * S;
  const int64_t n = i / per, r = i - n * per;
  const float* b = x + n * (int64_t)T * S + r;
  y[i] = b[S] - b[0];
}
void launch_delta_t(const float* x, float* y, int N, int T, int64_t S, cudaStream_t s) {
  const int64_t total = (int64_t)N * (T - 1) * S;
  if (total <= 0) return;
  delta_t_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, T, S, total);
  count_launch();
}

// fusion_module.py:67-73: F.interpolate(mode="linear", align_corners=False) along tokens.
__global__ void lerp_tokens_kernel(const float* x, float* y, int Tin, int Tout, int D, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = (int)(i % D);
  const int64_t r = i / D;
  const int t = (int)(r % Tout);
  const int64_t n = r / Tout;
  const float scale = (float)Tin / (float)Tout;
  float src = ((float)t + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  int i0 = (int)floorf(src);
  i0 = i0 > Tin - 1 ? Tin - 1 : i0;
  const int i1 = i0 + 1 > Tin - 1 ? Tin - 1 : i0 + 1;
  const float w1 = src - (float)i0;
  const float* b = x + n * (int64_t)Tin * D + d;
  y[i] = b[(int64_t)i0 * D] * (1.0f - w1) + b[(int64_t)i1 * D] * w1;
}
void launch_lerp_tokens(const float* x, float* y, int N, int Tin, int Tout, int D, cudaStream_t s) {
  const int64_t total = (int64_t)N * Tout * D;
  if (total == 0) return;
  lerp_tokens_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, y, Tin, Tout, D, total);
  count_launch();
}

__global__ void set_cls_kernel(const float* cls, float* y, int tokens, int D, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = (int)(i % D);
  const int64_t n = i / D;
  y[n * (int64_t)tokens * D + d] = cls[d];
}
void launch_set_cls(const float* cls, float* y, int N, int tokens, int D, cudaStream_t s) {
  const int64_t total = (int64_t)N * D;
  if (total == 0) return;
  set_cls_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cls, y, tokens, D, total);
  count_launch();
}

__global__ void copy_rows_kernel(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int width, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % width);
  const int64_t r = i / width;
  dst[r * dst_ld + c] = src[r * src_ld + c];
}
void launch_copy_rows(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int rows, int width, cudaStream_t s) {
  const int64_t total = (int64_t)rows * width;
  if (total == 0) return;
  copy_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, src_ld, dst, dst_ld, width, total);
  count_launch();
}

__global__ void fill_zero_kernel(float* x, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = 0.f;
}
void launch_fill_zero(float* x, int64_t n, cudaStream_t s) {
  if (n <= 0) return;
  fill_zero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n);
  count_launch();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// nn.LayerNorm (biased variance, eps 1e-5): one warp per row, two-pass in registers/L1.
__global__ void layernorm_kernel(const float* x, int64_t x_ld, const float* g, const float* b, float* y, int64_t y_ld, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * x_ld;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += xr[d];
  const float mu = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int d = lane; d < D; d += 32) { const float t = xr[d] - mu; v += t * t; }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)D + 1e-5f);
  float* yr = y + (int64_t)row * y_ld;
  for (int d = lane; d < D; d += 32) yr[d] = (xr[d] - mu) * rstd * g[d] + b[d];
}
void launch_layernorm(const float* x, int64_t x_ld, const float* g, const float* b, float* y, int64_t y_ld, int rows, int D, cudaStream_t s) {
  if (rows == 0) return;
  layernorm_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, x_ld, g, b, y, y_ld, rows, D);
  count_launch();
}

// Attention core of nn.MultiheadAttention (head dim 32 == warp width): one block per (window, head),
// K/V of the head staged in shared memory, one warp per query row, fp32 softmax.
__global__ void mha_core_kernel(const float* q, int q_ld, const float* k, int k_ld, const float* v, int v_ld,
                                float* o, int o_ld, int Tq, int Tk, int heads) {
  extern __shared__ float sm[];
  float* Ks = sm;                 // [Tk][33]
  float* Vs = sm + (size_t)Tk * 33;  // [Tk][33]
  const int n = blockIdx.x / heads, h = blockIdx.x % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < Tk; j += nwarps) {
    Ks[j * 33 + lane] = k[((int64_t)n * Tk + j) * k_ld + h * 32 + lane];
    Vs[j * 33 + lane] = v[((int64_t)n * Tk + j) * v_ld + h * 32 + lane];
  }
  __syncthreads();
  const float scale = 0.17677669529663688f;  // 1/sqrt(32)
  for (int i = warp; i < Tq; i += nwarps) {
    const float qd = q[((int64_t)n * Tq + i) * q_ld + h * 32 + lane] * scale;
    float acc = 0.f, mx = -INFINITY, den = 0.f;
    // pass 1: row max
    for (int j0 = 0; j0 < Tk; j0 += 32) {
      const int j = j0 + lane;
      float sc = -INFINITY;
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        const float qv = __shfl_sync(0xffffffffu, qd, d);
        if (j < Tk) dot = fmaf(qv, Ks[j * 33 + d], dot);
      }
      if (j < Tk) sc = dot;
      mx = fmaxf(mx, warp_max(sc));
    }
    // pass 2: exp, sum, PV
    for (int j0 = 0; j0 < Tk; j0 += 32) {
      const int j = j0 + lane;
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        const float qv = __shfl_sync(0xffffffffu, qd, d);
        if (j < Tk) dot = fmaf(qv, Ks[j * 33 + d], dot);
      }
      const float pj = (j < Tk) ? expf(dot - mx) : 0.f;
      den += warp_sum(pj);
      const int cnt = min(32, Tk - j0);
      for (int jj = 0; jj < cnt; ++jj) {
        const float pv = __shfl_sync(0xffffffffu, pj, jj);
        acc = fmaf(pv, Vs[(j0 + jj) * 33 + lane], acc);
      }
    }
    o[((int64_t)n * Tq + i) * o_ld + h * 32 + lane] = acc / den;
  }
}
void launch_mha_core(const float* q, int q_ld, const float* k, int k_ld, const float* v, int v_ld, float* o, int o_ld,
                     int N, int Tq, int Tk, int heads, cudaStream_t s) {
  if (N == 0) return;
  const size_t smem = (size_t)Tk * 33 * 2 * sizeof(float);
  mha_core_kernel<<<N * heads, 128, smem, s>>>(q, q_ld, k, k_ld, v, v_ld, o, o_ld, Tq, Tk, heads);
  count_launch();
}

// fusion_module.py:84-86: g = sigmoid(Linear(256,1)(h)); fused = g*v + (1-g)*a.  One warp per token.
__global__ void gate_blend_kernel(const float* h, const float* w2, const float* b2, const float* v, int v_ld,
                                  const float* a, int a_ld, float* out, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(h[(int64_t)row * D + d], w2[d], s);
  s = warp_sum(s) + b2[0];
  const float g = 1.0f / (1.0f + expf(-s));
  for (int d = lane; d < D; d += 32) {
    const float vv = v[(int64_t)row * v_ld + d], aa = a[(int64_t)row * a_ld + d];
    out[(int64_t)row * D + d] = g * vv + (1.0f - g) * aa;
  }
}
void launch_gate_blend(const float* h, const float* w2, const float* b2, const float* v, int v_ld, const float* a, int a_ld,
                       float* out, int rows, int D, cudaStream_t s) {
  if (rows == 0) return;
  gate_blend_kernel<<<(rows + 7) / 8, 256, 0, s>>>(h, w2, b2, v, v_ld, a, a_ld, out, rows, D);
  count_launch();
}

// classifier.py:17-19,34: LayerNorm(128) then Linear(128,1), squeeze -> one fp32 logit per window.
__global__ void ln_dot_kernel(const float* x, const float* g, const float* b, const float* w, const float* bias, float* out, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += xr[d];
  const float mu = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int d = lane; d < D; d += 32) { const float t = xr[d] - mu; v += t * t; }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)D + 1e-5f);
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) acc = fmaf((xr[d] - mu) * rstd * g[d] + b[d], w[d], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc + bias[0];
}
void launch_ln_dot(const float* x, const float* g, const float* b, const float* w, const float* bias, float* out, int rows, int D, cudaStream_t s) {
  if (rows == 0) return;
  ln_dot_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, g, b, w, bias, out, rows, D);
  count_launch();
}

}  // namespace lsd
