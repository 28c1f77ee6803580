// Log-mel front end: replaces librosa.feature.melspectrogram + power_to_db(ref=np.max) of the reference
// (app/preprocessing/audio.py:80-91; semantics restated in SURVEY.md App. D):
//   n_fft = win = 400 (periodic Hann), hop 160, center=True with zero padding, power 2, 80 Slaney mel
//   bands (0..8000 Hz, area-normalised), dB relative to the clip maximum, floored at -80 dB.
// Pass 1: one block per STFT frame — windowed frame staged in shared memory, 201-bin DFT from a
// shared twiddle table, sparse mel projection, block max -> atomicMax on the clip maximum.
// Pass 2: elementwise dB conversion against the clip maximum.
#include "lsd_kernels.h"
#include "umma.cuh"

#include <math.h>

namespace lsd {

constexpr int NFFT = 400, HOP = 160, NBINS = 201, NMEL = 80, MELW_MAX = 32;

__global__ void __launch_bounds__(256) logmel_power_kernel(const float* __restrict__ pcm, int64_t n_samples, int frames,
                                                           const float* __restrict__ hann, const float* __restrict__ tc,
                                                           const float* __restrict__ ts, const float* __restrict__ melw,
                                                           const int* __restrict__ mel_lo, const int* __restrict__ mel_cnt,
                                                           float* __restrict__ mel_power, float* clip_max) {
  __shared__ float xs[NFFT], cs[NFFT], sn[NFFT], pw[NBINS + 3], red[8];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int64_t start = (int64_t)f * HOP - NFFT / 2;
  for (int n = tid; n < NFFT; n += 256) {
    const int64_t i = start + n;
    xs[n] = (i >= 0 && i < n_samples) ? pcm[i] * hann[n] : 0.f;
    cs[n] = tc[n];
    sn[n] = ts[n];
  }
  __syncthreads();
  if (tid < NBINS) {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int n = 0; n < NFFT; ++n) {
      const float x = xs[n];
      re = fmaf(x, cs[idx], re);
      im = fmaf(x, sn[idx], im);
      idx += tid;
      if (idx >= NFFT) idx -= NFFT;
    }
    pw[tid] = re * re + im * im;
  }
  __syncthreads();
  float mx = 0.f;
  if (tid < NMEL) {
    const int lo = mel_lo[tid], cnt = mel_cnt[tid];
    float acc = 0.f;
    for (int j = 0; j < cnt; ++j) acc = fmaf(melw[tid * MELW_MAX + j], pw[lo + j], acc);
    mel_power[(int64_t)tid * frames + f] = acc;
    mx = acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomicMax(reinterpret_cast<int*>(clip_max), __float_as_int(m));  // non-negative floats order like ints
  }
}

void launch_logmel_power(const float* pcm, int64_t n_samples, int frames, const float* hann, const float* twid_cos,
                         const float* twid_sin, const float* melw, const int* mel_lo, const int* mel_cnt,
                         float* mel_power, float* clip_max, cudaStream_t s) {
  if (frames <= 0) return;
  logmel_power_kernel<<<frames, 256, 0, s>>>(pcm, n_samples, frames, hann, twid_cos, twid_sin, melw, mel_lo, mel_cnt,
                                             mel_power, clip_max);
  count_launch();
}

// power_to_db(S, ref=max, amin=1e-10, top_db=80): 10*log10(max(S,amin)) - 10*log10(max(ref,amin)), floored at max-80 = -80.
__global__ void logmel_db_kernel(float* mel, int64_t n, const float* clip_max) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // librosa rounds 10*log10(S) and 10*log10(ref) separately before subtracting: keep the products un-fused
  const float ref = __fmul_rn(10.0f, log10f(fmaxf(clip_max[0], 1e-10f)));
  const float v = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(mel[i], 1e-10f))), ref);
  mel[i] = fmaxf(v, -80.0f);
}
void launch_logmel_db(float* mel, int64_t n, const float* clip_max, cudaStream_t s) {
  if (n <= 0) return;
  logmel_db_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mel, n, clip_max);
  count_launch();
}

}  // namespace lsd

// ================================================================================================
// FFT version (batched over clips): replaces the per-frame direct DFT above on the product path.
// ================================================================================================
// One CTA = LM_FB consecutive frames of one clip.  The PCM segment the frames share (LM_FB-1 hops + one window, 11 KB) is
// staged in shared memory once — by one bulk async copy (TMA engine) when the segment is interior and 16-byte aligned, by
// coalesced loads otherwise (clip edges are the zero padding of center=True).  Two real frames are packed into one complex
// 400-point FFT (frame a -> real part, frame b -> imaginary part), computed as 16 x 25 Cooley-Tukey:
//   step 1  25 x 16-point DFTs in registers (radix 4 x 4) over the windowed samples n = 25*n1 + n2, times W400^(k1*n2)
//   step 2  16 x 25-point DFTs in registers (radix 5 x 5)            -> Z[k1 + 16*k2]
//   step 3  untangle the two real spectra, |.|^2 for bins 0..200      -> shared memory
//   step 4  sparse Slaney mel projection (80 x <= 32 bins), clip maximum (atomicMax on non-negative floats)
// ~21 kFLOP per frame instead of 322 kFLOP for the direct DFT; the dB pass stays separate (ref = clip maximum).
namespace lsd {

constexpr int LM_FB = 16, LM_PAIRS = LM_FB / 2;
constexpr int LM_SEG = (LM_FB - 1) * HOP + NFFT;       // 2800 samples
constexpr int LM_SEG_PAD = 2816;
constexpr int LM_PSTRIDE = 208;                          // power spectrum row stride (201 bins padded)
__constant__ float2 c_w16[16], c_w25[25], c_w5[5];

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmuli_neg(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = cmuli_neg(csub(a1, a3));
  a0 = cadd(s02, s13); a1 = cadd(d02, d13); a2 = csub(s02, s13); a3 = csub(d02, d13);
}
__device__ __forceinline__ void dft5(const float2* in, float2* out) {   // out[c] = sum_a in[a] * W5^(a*c)
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    float2 acc = in[0];
#pragma unroll
    for (int a = 1; a < 5; ++a) acc = cadd(acc, c == 0 ? in[a] : cmul(in[a], c_w5[(a * c) % 5]));
    out[c] = acc;
  }
}

__global__ void __launch_bounds__(256) logmel_fft_kernel(const float* __restrict__ pcm, const LmClip* __restrict__ clips, int n_clips,
                                                         const float* __restrict__ hann, const float2* __restrict__ w400,
                                                         const float* __restrict__ melw, const int* __restrict__ mel_lo,
                                                         const int* __restrict__ mel_cnt, float* __restrict__ mel_out,
                                                         float* __restrict__ clip_max) {
  extern __shared__ __align__(16) float lm_smem[];
  float* seg = lm_smem;                                              // LM_SEG_PAD floats
  float2* bufA = reinterpret_cast<float2*>(seg + LM_SEG_PAD);        // LM_PAIRS x 400
  float2* bufB = bufA + LM_PAIRS * NFFT;                             // LM_PAIRS x 400
  float* pw = reinterpret_cast<float*>(bufB);                        // LM_FB x LM_PSTRIDE (after step 2, bufB is dead)
  __shared__ uint64_t bar;
  __shared__ float red[8];
  const int tid = threadIdx.x;
  // clip of this block: last clip whose first block index is <= blockIdx.x
  int lo = 0, hi = n_clips - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (clips[mid].block0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const LmClip cl = clips[lo];
  const int f0 = ((int)blockIdx.x - cl.block0) * LM_FB;
  const long long s0 = (long long)f0 * HOP - NFFT / 2;              // clip sample of seg[0]
  const float* src = pcm + cl.pcm_off;
  const bool bulk = s0 >= 0 && s0 + LM_SEG <= cl.n_samples && ((reinterpret_cast<uintptr_t>(src + s0) & 15) == 0);
  if (bulk) {
    if (tid == 0) {
      umma::mbar_init(&bar, 1);
      umma::fence_barrier_init();
      umma::mbar_arrive_expect_tx(&bar, LM_SEG * 4);
      umma::bulk_g2s(seg, src + s0, LM_SEG * 4, &bar);
    }
    __syncthreads();
    umma::mbar_wait(&bar, 0);
  } else {
    for (int i = tid; i < LM_SEG; i += 256) {
      const long long s = s0 + i;
      seg[i] = (s >= 0 && s < cl.n_samples) ? src[s] : 0.f;          // center=True, pad_mode="constant"
    }
    __syncthreads();
  }
  // ---- step 1: 16-point DFTs over n1 (n = 25*n1 + n2), twiddle W400^(k1*n2) -> bufB[pair][k1*25 + n2]
  if (tid < LM_PAIRS * 25) {
    const int pr = tid / 25, n2 = tid - pr * 25;
    const float* xa = seg + (2 * pr) * HOP;
    const float* xb = xa + HOP;
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int n = 25 * n1 + n2;
      const float h = hann[n];
      v[n1] = make_float2(h * xa[n], h * xb[n]);
    }
    // T[b][c] = sum_a v[4a+b] W4^(a*c): in place on (v[b], v[4+b], v[8+b], v[12+b]) -> index 4c + b
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // Y[c + 4d] = sum_b W4^(b*d) * (W16^(b*c) * T[b][c])
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float2 t0 = v[4 * c], t1 = v[4 * c + 1], t2 = v[4 * c + 2], t3 = v[4 * c + 3];
      if (c > 0) { t1 = cmul(t1, c_w16[c]); t2 = cmul(t2, c_w16[2 * c]); t3 = cmul(t3, c_w16[3 * c]); }
      dft4(t0, t1, t2, t3);                                           // t_d = Y[c + 4d]
      float2* dst = bufB + pr * NFFT + n2;
      dst[(c) * 25] = cmul(t0, w400[c * n2]);
      dst[(c + 4) * 25] = cmul(t1, w400[(c + 4) * n2]);
      dst[(c + 8) * 25] = cmul(t2, w400[(c + 8) * n2]);
      dst[(c + 12) * 25] = cmul(t3, w400[(c + 12) * n2]);
    }
  }
  __syncthreads();
  // ---- step 2: 25-point DFTs over n2 for every k1 -> bufA[pair][k1 + 16*k2]
  if (tid < LM_PAIRS * 16) {
    const int pr = tid >> 4, k1 = tid & 15;
    const float2* u = bufB + pr * NFFT + k1 * 25;
    float2 t[25];                                                     // t[5c + b] = W25^(b*c) * sum_a u[5a+b] W5^(a*c)
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      float2 in[5], out[5];
#pragma unroll
      for (int a = 0; a < 5; ++a) in[a] = u[5 * a + b];
      dft5(in, out);
#pragma unroll
      for (int c = 0; c < 5; ++c) t[5 * c + b] = (b == 0 || c == 0) ? out[c] : cmul(out[c], c_w25[b * c]);
    }
    float2* z = bufA + pr * NFFT + k1;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      float2 out[5];
      dft5(&t[5 * c], out);                                           // out[d] = Z[k1 + 16*(c + 5d)]
#pragma unroll
      for (int d = 0; d < 5; ++d) z[16 * (c + 5 * d)] = out[d];
    }
  }
  __syncthreads();
  // ---- step 3: untangle the two real spectra and take |.|^2
  for (int i = tid; i < LM_PAIRS * NBINS; i += 256) {
    const int pr = i / NBINS, k = i - pr * NBINS;
    const float2 za = bufA[pr * NFFT + k], zc = bufA[pr * NFFT + (k == 0 ? 0 : NFFT - k)];
    const float ar = 0.5f * (za.x + zc.x), ai = 0.5f * (za.y - zc.y);      // X_a = (Z[k] + conj(Z[N-k])) / 2
    const float br = 0.5f * (za.y + zc.y), bi = -0.5f * (za.x - zc.x);     // X_b = (Z[k] - conj(Z[N-k])) / (2i)
    pw[(2 * pr) * LM_PSTRIDE + k] = ar * ar + ai * ai;
    pw[(2 * pr + 1) * LM_PSTRIDE + k] = br * br + bi * bi;
  }
  __syncthreads();
  // ---- step 4: mel projection + clip maximum
  float mx = 0.f;
  for (int i = tid; i < LM_FB * NMEL; i += 256) {
    const int m = i / LM_FB, f = i - m * LM_FB;                            // frame fastest: coalesced rows of the (80, frames) output
    if (f0 + f < cl.frames) {
      const int lo_b = mel_lo[m], cnt = mel_cnt[m];
      const float* p = pw + f * LM_PSTRIDE + lo_b;
      float acc = 0.f;
      for (int j = 0; j < cnt; ++j) acc = fmaf(melw[m * MELW_MAX + j], p[j], acc);
      mel_out[cl.mel_off + (long long)m * cl.frames + f0 + f] = acc;
      mx = fmaxf(mx, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomicMax(reinterpret_cast<int*>(clip_max + lo), __float_as_int(m));  // non-negative floats order like ints
  }
}

__global__ void logmel_db_batched_kernel(float* __restrict__ mel, const LmClip* __restrict__ clips, const float* __restrict__ clip_max) {
  const LmClip cl = clips[blockIdx.y];
  const long long n = (long long)NMEL * cl.frames;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // librosa rounds 10*log10(S) and 10*log10(ref) separately before subtracting: keep the products un-fused
  const float ref = __fmul_rn(10.0f, log10f(fmaxf(clip_max[blockIdx.y], 1e-10f)));
  float* p = mel + cl.mel_off + i;
  const float v = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(*p, 1e-10f))), ref);
  *p = fmaxf(v, -80.0f);
}

size_t logmel_fft_smem_bytes() { return (size_t)LM_SEG_PAD * 4 + (size_t)2 * LM_PAIRS * NFFT * sizeof(float2); }
int logmel_frames_per_block() { return LM_FB; }

// __constant__ twiddles and function attributes are per device: called from lsd_create with the handle's device current
void init_logmel_fft_constants() {
  float2 w16[16], w25[25], w5[5];
  const double PI = 3.14159265358979323846;
  for (int k = 0; k < 16; ++k) w16[k] = make_float2((float)cos(2.0 * PI * k / 16), (float)-sin(2.0 * PI * k / 16));
  for (int k = 0; k < 25; ++k) w25[k] = make_float2((float)cos(2.0 * PI * k / 25), (float)-sin(2.0 * PI * k / 25));
  for (int k = 0; k < 5; ++k) w5[k] = make_float2((float)cos(2.0 * PI * k / 5), (float)-sin(2.0 * PI * k / 5));
  cudaMemcpyToSymbol(c_w16, w16, sizeof(w16));
  cudaMemcpyToSymbol(c_w25, w25, sizeof(w25));
  cudaMemcpyToSymbol(c_w5, w5, sizeof(w5));
  cudaFuncSetAttribute(logmel_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)logmel_fft_smem_bytes());
}

void launch_logmel_fft(const float* pcm, const LmClip* clips, int n_clips, int total_blocks, int max_frames, const float* hann,
                       const float2* w400, const float* melw, const int* mel_lo, const int* mel_cnt, float* mel_out, float* clip_max,
                       cudaStream_t s) {
  if (total_blocks <= 0) return;
  logmel_fft_kernel<<<total_blocks, 256, logmel_fft_smem_bytes(), s>>>(pcm, clips, n_clips, hann, w400, melw, mel_lo, mel_cnt, mel_out, clip_max);
  count_launch();
  const long long n = (long long)NMEL * max_frames;
  logmel_db_batched_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n_clips), 256, 0, s>>>(mel_out, clips, clip_max);
  count_launch();
}

}  // namespace lsd
