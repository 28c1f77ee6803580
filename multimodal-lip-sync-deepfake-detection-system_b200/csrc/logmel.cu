// Log-mel front end: replaces librosa.feature.melspectrogram + power_to_db(ref=np.max) of the reference
// (app/preprocessing/audio.py:80-91; semantics restated in SURVEY.md App. D):
//   n_fft = win = 400 (periodic Hann), hop 160, center=True with zero padding, power 2, 80 Slaney mel
//   bands (0..8000 Hz, area-normalised), dB relative to the clip maximum, floored at -80 dB.
// Pass 1: one block per STFT frame — windowed frame staged in shared memory, 201-bin DFT from a
// shared twiddle table, sparse mel projection, block max -> atomicMax on the clip maximum.
// Pass 2: elementwise dB conversion against the clip maximum.
#include "lsd_kernels.h"

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
