// Host side of the log-mel front end: Slaney mel filterbank / Hann / twiddle tables and the lsd_logmel entry point.
// Restates librosa>=0.10 `filters.mel(sr=16000, n_fft=400, n_mels=80, htk=False, norm="slaney")` as called by
// app/preprocessing/audio.py:80-88 of the reference (librosa itself is an un-vendored dependency).
#include "../../include/lsd_b200.h"
#include "lsd_internal.h"
#include "lsd_kernels.h"

#include <cmath>
#include <cstring>

namespace {
constexpr int NFFT = 400, NBINS = 201, NMEL = 80, MELW_MAX = 32;
constexpr double SR = 16000.0, FMIN = 0.0, FMAX = 8000.0;

double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}
}  // namespace

int init_logmel_tables(lsd_handle* h) {
  std::vector<float> hann(NFFT), tc(NFFT), ts(NFFT), melw((size_t)NMEL * MELW_MAX, 0.f);
  std::vector<int> lo(NMEL, 0), cnt(NMEL, 0);
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < NFFT; ++n) {
    hann[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / NFFT));  // periodic Hann (fftbins=True)
    tc[n] = (float)std::cos(2.0 * PI * n / NFFT);
    ts[n] = (float)(-std::sin(2.0 * PI * n / NFFT));
  }
  double mel_f[NMEL + 2];
  const double m_lo = hz_to_mel(FMIN), m_hi = hz_to_mel(FMAX);
  for (int i = 0; i < NMEL + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (NMEL + 1));
  for (int i = 0; i < NMEL; ++i) {
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    int first = -1, last = -1;
    float row[NBINS];
    for (int k = 0; k < NBINS; ++k) {
      const double fk = (SR / 2.0) * k / (NBINS - 1);
      const double lower = (fk - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
      const double upper = (mel_f[i + 2] - fk) / (mel_f[i + 2] - mel_f[i + 1]);
      const float tri = (float)std::fmax(0.0, std::fmin(lower, upper));
      row[k] = (float)((double)tri * enorm);
      if (row[k] != 0.f) { if (first < 0) first = k; last = k; }
    }
    if (first < 0) { first = 0; last = -1; }
    if (last - first + 1 > MELW_MAX) return lsd_fail(h, LSD_ERR_UNSUPPORTED, "mel filter %d wider than %d bins", i, MELW_MAX);
    lo[i] = first; cnt[i] = last - first + 1;
    for (int k = first; k <= last; ++k) melw[(size_t)i * MELW_MAX + (k - first)] = row[k];
  }
  const size_t nf = 3 * NFFT + (size_t)NMEL * MELW_MAX + 2 * NFFT;   // ... + interleaved (cos, -sin) twiddles of the FFT kernel
  const size_t bytes = nf * sizeof(float) + 2 * NMEL * sizeof(int);
  cudaError_t e = cudaSetDevice(h->device);
  if (e == cudaSuccess) e = cudaMalloc(&h->mel_tables, bytes);
  if (e != cudaSuccess) return lsd_fail(h, LSD_ERR_CUDA, "log-mel tables: %s", cudaGetErrorString(e));
  std::vector<char> host(bytes);
  float* f = reinterpret_cast<float*>(host.data());
  memcpy(f, hann.data(), NFFT * 4); memcpy(f + NFFT, tc.data(), NFFT * 4); memcpy(f + 2 * NFFT, ts.data(), NFFT * 4);
  memcpy(f + 3 * NFFT, melw.data(), melw.size() * 4);
  for (int n = 0; n < NFFT; ++n) { f[3 * NFFT + melw.size() + 2 * n] = tc[n]; f[3 * NFFT + melw.size() + 2 * n + 1] = ts[n]; }
  int* ip = reinterpret_cast<int*>(f + nf);
  memcpy(ip, lo.data(), NMEL * 4); memcpy(ip + NMEL, cnt.data(), NMEL * 4);
  e = cudaMemcpy(h->mel_tables, host.data(), bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return lsd_fail(h, LSD_ERR_CUDA, "log-mel tables copy: %s", cudaGetErrorString(e));
  const float* d = reinterpret_cast<const float*>(h->mel_tables);
  h->d_hann = d; h->d_cos = d + NFFT; h->d_sin = d + 2 * NFFT; h->d_melw = d + 3 * NFFT;
  h->d_w400 = d + 3 * NFFT + (size_t)NMEL * MELW_MAX;
  h->d_mel_lo = reinterpret_cast<const int*>(d + nf); h->d_mel_cnt = h->d_mel_lo + NMEL;
  lsd::init_logmel_fft_constants();
  return 0;
}

extern "C" int lsd_logmel_frames(int64_t n_samples) { return n_samples < 0 ? 0 : (int)(1 + n_samples / 160); }

extern "C" int lsd_logmel(lsd_handle* h, const float* pcm, const int64_t* clip_offsets_host, int n_clips, float* mel_out,
                          const int64_t* mel_offsets_host, float* scratch, void* stream) {
  if (!h) return LSD_ERR_ARG;
  if (n_clips == 0) return LSD_OK;
  if (!pcm || !clip_offsets_host || !mel_out || !mel_offsets_host || !scratch) return lsd_fail(h, LSD_ERR_ARG, "lsd_logmel: null pointer argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev != h->device ? prev_dev : -1};
  cudaError_t e = prev_dev != h->device ? cudaSetDevice(h->device) : cudaSuccess;
  if (e == cudaSuccess) e = cudaMemsetAsync(scratch, 0, sizeof(float) * n_clips, st);
  if (e != cudaSuccess) return lsd_fail(h, LSD_ERR_CUDA, "lsd_logmel: %s", cudaGetErrorString(e));
  if (n_clips > 65535) return lsd_fail(h, LSD_ERR_SHAPE, "lsd_logmel: at most 65535 clips per call");
  // clip table (device): built on the host, uploaded on the stream; the device buffer is owned by the handle and grows on demand
  const int fb = lsd::logmel_frames_per_block();
  h->lm_clips_host.resize((size_t)n_clips * sizeof(lsd::LmClip));
  lsd::LmClip* tab = reinterpret_cast<lsd::LmClip*>(h->lm_clips_host.data());
  long long blocks = 0;
  int max_frames = 0;
  for (int c = 0; c < n_clips; ++c) {
    const int64_t n = clip_offsets_host[c + 1] - clip_offsets_host[c];
    if (n < 0) return lsd_fail(h, LSD_ERR_SHAPE, "lsd_logmel: clip %d has negative length", c);
    const int frames = lsd_logmel_frames(n);
    tab[c].pcm_off = clip_offsets_host[c]; tab[c].n_samples = n; tab[c].mel_off = mel_offsets_host[c];
    tab[c].frames = frames; tab[c].block0 = (int)blocks;
    blocks += (frames + fb - 1) / fb;
    if (frames > max_frames) max_frames = frames;
  }
  if (blocks > 0x7fffffffLL) return lsd_fail(h, LSD_ERR_SHAPE, "lsd_logmel: too many frames in one call");
  if (h->lm_clips_cap < h->lm_clips_host.size()) {
    if (h->lm_clips) { cudaStreamSynchronize(st); cudaFree(h->lm_clips); h->lm_clips = nullptr; }
    h->lm_clips_cap = h->lm_clips_host.size() * 2;
    if (cudaMalloc(&h->lm_clips, h->lm_clips_cap) != cudaSuccess) { h->lm_clips_cap = 0; return lsd_fail(h, LSD_ERR_CUDA, "lsd_logmel: clip table allocation failed"); }
  }
  // the clip table is shared by all calls on this handle: the upload waits for the kernels of the previous call (which may have
  // run on another stream) before it overwrites the table
  if (!h->ev_lm_clips && cudaEventCreateWithFlags(&h->ev_lm_clips, cudaEventDisableTiming) != cudaSuccess)
    return lsd_fail(h, LSD_ERR_CUDA, "lsd_logmel: event creation failed");
  if (h->lm_clips_used) cudaStreamWaitEvent(st, h->ev_lm_clips, 0);
  e = cudaMemcpyAsync(h->lm_clips, tab, h->lm_clips_host.size(), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return lsd_fail(h, LSD_ERR_CUDA, "lsd_logmel: %s", cudaGetErrorString(e));
  lsd::launch_logmel_fft(pcm, reinterpret_cast<const lsd::LmClip*>(h->lm_clips), n_clips, (int)blocks, max_frames, h->d_hann,
                         reinterpret_cast<const float2*>(h->d_w400), h->d_melw, h->d_mel_lo, h->d_mel_cnt, mel_out, scratch, st);
  cudaEventRecord(h->ev_lm_clips, st);
  h->lm_clips_used = true;
  e = cudaGetLastError();
  if (e != cudaSuccess) return lsd_fail(h, LSD_ERR_CUDA, "lsd_logmel: %s", cudaGetErrorString(e));
  return LSD_OK;
}
