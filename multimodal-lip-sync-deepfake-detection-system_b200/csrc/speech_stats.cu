// Per-window speaking-alignment score and mouth-motion statistics on the device.
//
// Replaces, for the windows of a track, the per-window host numpy code of the reference:
//   Predictor._speaking_alignment_score    app/inference/predictor.py:333-370
//   Predictor._mouth_motion_energy_check   app/inference/predictor.py:374-419  (statistics; the thresholds stay on the host)
// which `_predict_long_video` calls once per window on decoded float32 crops (:793-800, :1118).  Sliding windows overlap
// (stride 8, length 32), so the frame-difference energies are computed ONCE per frame pair of the uint8 track
// (`track_motion_kernel`, HBM-bound: every frame is read twice, 2 x 27.6 KB per pair) and every window only combines its
// 31 values with its slice of the clip log-mel (`speech_stats_kernel`).
//
// Arithmetic follows numpy's: crops are uint8 / 255 in float32 (video.py:552-556); the channel mean is ((r + g) + b) / 3 in
// float32; the resampling (np.linspace / np.interp), the z-scores of the resampled motion and np.corrcoef run in float64 as in
// numpy; the audio energy (mean over F), its z-score (mean / std with numpy's pairwise float32 summation for <= 128
// elements) run in float32.  Large means (|diff| over H*W, the mel mean of the mouth check) accumulate in float64 and are
// rounded once — within 1e-7 relative of numpy's pairwise float32 sums.
#include "lsd_kernels.h"

#include <math.h>
#include <stdint.h>

namespace lsd {

__device__ __forceinline__ float ss_gray_u8(const uint8_t* p) {
  const float r = (float)p[0] / 255.0f, g = (float)p[1] / 255.0f, b = (float)p[2] / 255.0f;
  return __fdiv_rn(__fadd_rn(__fadd_rn(r, g), b), 3.0f);
}

// One block per frame pair f: sum |gray[f+1] - gray[f]| over all pixels and over the lower half rows (h >= H/2).
// LAYOUT 1: uint8 (n_frames, H, W, 3) track; LAYOUT 0: float32 (3, T, H, W) window (the reference's visual_np).
template <int LAYOUT>
__global__ void __launch_bounds__(256) track_motion_kernel(const void* __restrict__ video, int n_frames, int H, int W,
                                                           float* __restrict__ motion_full, float* __restrict__ motion_low) {
  __shared__ double red[2][8];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int npix = H * W, low0 = (H / 2) * W;
  double s_all = 0.0, s_low = 0.0;
  for (int i = tid; i < npix; i += 256) {
    float g0, g1;
    if (LAYOUT == 1) {
      const uint8_t* v = reinterpret_cast<const uint8_t*>(video);
      g0 = ss_gray_u8(v + ((size_t)f * npix + i) * 3);
      g1 = ss_gray_u8(v + ((size_t)(f + 1) * npix + i) * 3);
    } else {
      const float* v = reinterpret_cast<const float*>(video);
      const size_t cs = (size_t)n_frames * npix;
      const size_t o0 = (size_t)f * npix + i, o1 = o0 + npix;
      g0 = __fdiv_rn(__fadd_rn(__fadd_rn(v[o0], v[o0 + cs]), v[o0 + 2 * cs]), 3.0f);
      g1 = __fdiv_rn(__fadd_rn(__fadd_rn(v[o1], v[o1 + cs]), v[o1 + 2 * cs]), 3.0f);
    }
    const float d = fabsf(__fsub_rn(g1, g0));
    s_all += (double)d;
    if (i >= low0) s_low += (double)d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_all += __shfl_xor_sync(0xffffffffu, s_all, o);
    s_low += __shfl_xor_sync(0xffffffffu, s_low, o);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = s_all; red[1][tid >> 5] = s_low; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, l = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; l += red[1][i]; }
    motion_full[f] = (float)(a / (double)npix);
    motion_low[f] = (float)(l / (double)(npix - low0));
  }
}

void launch_track_motion(const void* video, int layout, int n_frames, int H, int W, float* motion_full, float* motion_low,
                         cudaStream_t s) {
  if (n_frames < 2) return;
  if (layout == 1) track_motion_kernel<1><<<n_frames - 1, 256, 0, s>>>(video, n_frames, H, W, motion_full, motion_low);
  else track_motion_kernel<0><<<n_frames - 1, 256, 0, s>>>(video, n_frames, H, W, motion_full, motion_low);
  count_launch();
}

// numpy's pairwise float32 sum for n <= 128 (one block of PW_BLOCKSIZE): eight accumulators over strides of 8, combined as
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), then the tail sequentially; plain sequential sum below 8 elements.
__device__ float ss_pairwise_f32(const float* x, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, x[i]);
    return r;
  }
  float r[8];
  for (int k = 0; k < 8; ++k) r[k] = x[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], x[i + k]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, x[i]);
  return res;
}

constexpr int SS_MAX = 128;   // T <= 128 frames and Ta <= 128 mel columns per window

// One block (128 threads) per window.
__global__ void __launch_bounds__(128) speech_stats_kernel(const float* __restrict__ motion_full, const float* __restrict__ motion_low,
                                                           const int32_t* __restrict__ v_starts, const int32_t* __restrict__ a_starts,
                                                           int T, const float* __restrict__ mel, int F, int Ta_full, int Ta,
                                                           float* __restrict__ score, float* __restrict__ mouth_motion,
                                                           float* __restrict__ audio_energy) {
  __shared__ float a_e[SS_MAX], sq[SS_MAX];
  __shared__ double m_r[SS_MAX], part[4];
  __shared__ float mot[SS_MAX];
  const int w = blockIdx.x, tid = threadIdx.x;
  const int vs = v_starts[w], as = a_starts[w];
  // audio energy per column: mean over F (float32, sequential over f like numpy's axis-0 reduction); overall sum in float64
  double col_sum = 0.0;
  if (tid < Ta) {
    int col = as + tid;
    col = col >= Ta_full ? Ta_full - 1 : col;            // _align_audio_chunk: pad by repeating the last column
    float s = 0.f;
    for (int f = 0; f < F; ++f) { const float v = mel[(size_t)f * Ta_full + col]; s = __fadd_rn(s, v); col_sum += (double)v; }
    a_e[tid] = __fdiv_rn(s, (float)F);
  }
  if (tid < T) mot[tid] = tid == 0 ? motion_full[vs] : motion_full[vs + tid - 1];   // concatenate([motion[:1], motion])
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) col_sum += __shfl_xor_sync(0xffffffffu, col_sum, o);
  if ((tid & 31) == 0) part[tid >> 5] = col_sum;
  __syncthreads();
  // np.interp(linspace(0,1,Ta), linspace(0,1,T), motion) in float64
  if (tid < Ta && T >= 2) {
    const double step_old = 1.0 / (double)(T - 1), step_new = Ta > 1 ? 1.0 / (double)(Ta - 1) : 0.0;
    const double x = tid == Ta - 1 && Ta > 1 ? 1.0 : (double)tid * step_new;
    auto xo = [&](int i) { return i == T - 1 ? 1.0 : (double)i * step_old; };
    double val;
    if (x >= xo(T - 1)) val = (double)mot[T - 1];
    else {
      int j = (int)floor(x / step_old);
      j = j > T - 2 ? T - 2 : (j < 0 ? 0 : j);
      while (j > 0 && xo(j) > x) --j;                      // guard the floor() against rounding of x / step
      while (j < T - 2 && xo(j + 1) <= x) ++j;
      const double slope = ((double)mot[j + 1] - (double)mot[j]) / (xo(j + 1) - xo(j));
      val = slope * (x - xo(j)) + (double)mot[j];
    }
    m_r[tid] = val;
  }
  __syncthreads();
  if (tid == 0) {
    float sc = 0.5f;
    if (T >= 2 && Ta >= 2) {
      // z-score of the resampled motion (float64)
      double mu = 0.0;
      for (int i = 0; i < Ta; ++i) mu += m_r[i];
      mu /= (double)Ta;
      double var = 0.0;
      for (int i = 0; i < Ta; ++i) { const double d = m_r[i] - mu; var += d * d; }
      const double sig = sqrt(var / (double)Ta);
      double m_abs = 0.0;
      for (int i = 0; i < Ta; ++i) { m_r[i] = sig < 1e-6 ? 0.0 : (m_r[i] - mu) / sig; m_abs += fabs(m_r[i]); }
      // z-score of the audio energy (float32, numpy pairwise summation)
      const float amu = __fdiv_rn(ss_pairwise_f32(a_e, Ta), (float)Ta);
      for (int i = 0; i < Ta; ++i) { const float d = __fsub_rn(a_e[i], amu); sq[i] = __fmul_rn(d, d); }
      const float asig = sqrtf(__fdiv_rn(ss_pairwise_f32(sq, Ta), (float)Ta));
      double a_abs = 0.0;
      for (int i = 0; i < Ta; ++i) {
        a_e[i] = asig < 1e-6f ? 0.0f : __fdiv_rn(__fsub_rn(a_e[i], amu), asig);
        a_abs += fabs((double)a_e[i]);
      }
      if (m_abs >= 1e-6 && a_abs >= 1e-6) {
        // np.corrcoef(m, a)[0, 1] in float64
        double mm = 0.0, am = 0.0;
        for (int i = 0; i < Ta; ++i) { mm += m_r[i]; am += (double)a_e[i]; }
        mm /= (double)Ta; am /= (double)Ta;
        double cxy = 0.0, cxx = 0.0, cyy = 0.0;
        for (int i = 0; i < Ta; ++i) {
          const double dx = m_r[i] - mm, dy = (double)a_e[i] - am;
          cxy += dx * dy; cxx += dx * dx; cyy += dy * dy;
        }
        double corr = cxy / sqrt(cxx * cyy);
        if (corr == corr) {                                // not NaN
          corr = corr > 1.0 ? 1.0 : (corr < -1.0 ? -1.0 : corr);
          double v = (corr + 1.0) * 0.5;
          v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
          sc = (float)v;
        }
      }
    }
    score[w] = sc;
    // mouth-motion statistics (predictor.py:395-402)
    double ml = 0.0;
    for (int t = 0; t + 1 < T; ++t) ml += (double)motion_low[vs + t];
    mouth_motion[w] = T >= 2 ? (float)(ml / (double)(T - 1)) : 0.f;
    audio_energy[w] = T >= 2 ? (float)((part[0] + part[1] + part[2] + part[3]) / ((double)F * (double)Ta)) : 0.f;
  }
}

int speech_stats_max() { return SS_MAX; }

void launch_speech_stats(const float* motion_full, const float* motion_low, const int32_t* v_starts, const int32_t* a_starts, int n_windows,
                         int T, const float* mel, int F, int Ta_full, int Ta, float* score, float* mouth_motion, float* audio_energy,
                         cudaStream_t s) {
  if (n_windows <= 0) return;
  speech_stats_kernel<<<n_windows, 128, 0, s>>>(motion_full, motion_low, v_starts, a_starts, T, mel, F, Ta_full, Ta, score, mouth_motion,
                                                audio_energy);
  count_launch();
}

}  // namespace lsd

// ================================================================================================
// Energy VAD: the per-frame Python loops of detect_voice_activity
// (app/preprocessing/audio.py:178-230).  frame i = samples [160 i, min(160 i + 400, n)); energy = mean of squares; the mask is
// energy >= threshold, then OR-smoothed over [i-1, i+1].  The two order statistics behind the threshold (median, 20th
// percentile) are taken on the host from the energies, with numpy, exactly as the reference does.
// ================================================================================================
namespace lsd {

__global__ void __launch_bounds__(128) frame_energy_kernel(const float* __restrict__ pcm, long long n, int n_frames, float* __restrict__ energy) {
  // one warp per frame: 400 samples, float64 accumulation, rounded once (numpy: pairwise float32 mean of y*y)
  const int f = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= n_frames) return;
  const long long a = (long long)f * 160, b = a + 400 < n ? a + 400 : n;
  double s = 0.0;
  for (long long i = a + lane; i < b; i += 32) { const float v = pcm[i]; s += (double)__fmul_rn(v, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) energy[f] = (float)(s / (double)(b - a));
}

__global__ void vad_mask_kernel(const float* __restrict__ energy, int n_frames, float threshold, uint8_t* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames) return;
  bool v = energy[i] >= threshold;
  if (i > 0) v = v || energy[i - 1] >= threshold;
  if (i + 1 < n_frames) v = v || energy[i + 1] >= threshold;
  mask[i] = v ? 1 : 0;
}

void launch_frame_energy(const float* pcm, long long n, int n_frames, float* energy, cudaStream_t s) {
  if (n_frames <= 0) return;
  frame_energy_kernel<<<(n_frames + 3) / 4, 128, 0, s>>>(pcm, n, n_frames, energy);
  count_launch();
}
void launch_vad_mask(const float* energy, int n_frames, float threshold, uint8_t* mask, cudaStream_t s) {
  if (n_frames <= 0) return;
  vad_mask_kernel<<<(n_frames + 255) / 256, 256, 0, s>>>(energy, n_frames, threshold, mask);
  count_launch();
}

}  // namespace lsd
