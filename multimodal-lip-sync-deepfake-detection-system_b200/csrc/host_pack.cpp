// Host-side transport helper (no GPU work, plain C++ compiled by the host compiler): lossless fp32 -> uint8 packing of
// window pixels for the H2D copy.
//
// The reference builds its float windows from uint8 mouth crops as `astype(np.float32) / 255.0` (app/preprocessing/video.py:552-556),
// so every pixel of a window handed to `_run_chunked_inference` is exactly fl(k / 255.0f) for an integer k in [0, 255].  Such a
// window can cross PCIe as k (one byte instead of four): the device normalisation of uint8 input reproduces fl(k / 255.0f) bit
// for bit (umma_conv.cu: video_rows, tests/test_host_logic.py), so the logits do not change.  lsd_host_pack_u8_exact() verifies the
// property for every value (by redoing the reference's division) while it packs, and reports failure for anything else — the
// caller then ships the fp32 values as before.
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/lsd_b200.h"

namespace {

constexpr int64_t CHUNK = 1 << 18;   // elements per work item (1 MB of fp32)

bool pack_scalar(const float* s, uint8_t* d, int64_t n) {
  bool ok = true;
  for (int64_t i = 0; i < n; ++i) {
    const float x = s[i];
    const float t = x * 255.0f;
    int k = (t >= 0.0f && t <= 255.5f) ? (int)__builtin_lrintf(t) : 0;
    if (k > 255) k = 255;
    const float back = (float)k / 255.0f;
    ok &= (__builtin_memcmp(&back, &x, 4) == 0);   // bit pattern: false for NaN, -0.0 and anything that is not exactly fl(k/255)
    d[i] = (uint8_t)k;
  }
  return ok;
}

// fl(k / 255.0f) without the divider: q0 = k * fl(1/255), one Newton step on the exact residual.  Verified against the division for
// all 256 values of k before it is used (newton_ok()).
inline float div255_newton(float k) {
  const float r = 1.0f / 255.0f;
  const float q0 = k * r;
  const float e = __builtin_fmaf(-q0, 255.0f, k);
  return __builtin_fmaf(e, r, q0);
}
bool newton_ok() {
  for (int k = 0; k < 256; ++k) {
    const float a = div255_newton((float)k), b = (float)k / 255.0f;
    if (__builtin_memcmp(&a, &b, 4) != 0) return false;
  }
  return true;
}

template <bool NEWTON>
__attribute__((target("avx2,fma"))) bool pack_avx2(const float* s, uint8_t* d, int64_t n) {
  const __m256 c255 = _mm256_set1_ps(255.0f), rcp = _mm256_set1_ps(1.0f / 255.0f);
  const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  const __m256i hi = _mm256_set1_epi32(255);
  __m256i bad = _mm256_setzero_si256();
  int64_t i = 0;
  for (; i + 32 <= n; i += 32) {
    __m256i k[4];
#pragma GCC unroll 4
    for (int j = 0; j < 4; ++j) {
      const __m256 x = _mm256_loadu_ps(s + i + 8 * j);
      k[j] = _mm256_cvtps_epi32(_mm256_mul_ps(x, c255));                  // round to nearest (NaN / overflow -> INT_MIN)
      const __m256 kf = _mm256_cvtepi32_ps(k[j]);
      __m256 back;
      if (NEWTON) {
        const __m256 q0 = _mm256_mul_ps(kf, rcp);
        back = _mm256_fmadd_ps(_mm256_fnmadd_ps(q0, c255, kf), rcp, q0);
      } else {
        back = _mm256_div_ps(kf, c255);                                   // the reference's own operation
      }
      bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_castps_si256(back), _mm256_castps_si256(x)));   // bit patterns differ
      bad = _mm256_or_si256(bad, _mm256_cmpgt_epi32(k[j], hi));
      bad = _mm256_or_si256(bad, _mm256_cmpgt_epi32(_mm256_setzero_si256(), k[j]));
    }
    const __m256i p01 = _mm256_packus_epi32(k[0], k[1]);   // per 128-bit lane: k0[0:4] k1[0:4] | k0[4:8] k1[4:8]
    const __m256i p23 = _mm256_packus_epi32(k[2], k[3]);
    const __m256i p = _mm256_packus_epi16(p01, p23);       // lane 0: k0[0:4] k1[0:4] k2[0:4] k3[0:4], lane 1: the upper halves
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), _mm256_permutevar8x32_epi32(p, perm));
  }
  bool ok = _mm256_testz_si256(bad, bad) != 0;
  if (i < n) ok &= pack_scalar(s + i, d + i, n - i);
  return ok;
}

// The same check 64 values at a time (AVX-512: compare-into-mask and the down-converting store take the place of the three OR chains and
// the pack / permute sequence; measured 14 - 18 % faster per thread than the AVX2 loop).  Tails go through the AVX2 loop.
template <bool NEWTON>
__attribute__((target("avx512f,avx512bw,avx512vl,avx2,fma"))) bool pack_avx512(const float* s, uint8_t* d, int64_t n) {
  const __m512 c255 = _mm512_set1_ps(255.0f), rcp = _mm512_set1_ps(1.0f / 255.0f);
  const __m512i hi = _mm512_set1_epi32(255);
  __mmask16 bad = 0;
  int64_t i = 0;
  // 64-byte aligned destinations (the pinned staging buffers are) take non-temporal stores: no read-for-ownership of the staging
  // lines, i.e. 14 % fewer host-DRAM bytes per window (3.58 MB read + 0.9 MB written instead of + 0.9 MB read for ownership) — the
  // resource that bounds this route when every GPU of a box is fed at once
  const bool nt = (reinterpret_cast<uintptr_t>(d) & 63) == 0;
  for (; i + 64 <= n; i += 64) {
    __m128i b[4];
#pragma GCC unroll 4
    for (int j = 0; j < 4; ++j) {
      const __m512 x = _mm512_loadu_ps(s + i + 16 * j);
      const __m512i k = _mm512_cvtps_epi32(_mm512_mul_ps(x, c255));         // round to nearest (NaN / overflow -> INT_MIN)
      const __m512 kf = _mm512_cvtepi32_ps(k);
      __m512 back;
      if (NEWTON) {
        const __m512 q0 = _mm512_mul_ps(kf, rcp);
        back = _mm512_fmadd_ps(_mm512_fnmadd_ps(q0, c255, kf), rcp, q0);
      } else {
        back = _mm512_div_ps(kf, c255);                                     // the reference's own operation
      }
      bad |= _mm512_cmpneq_epi32_mask(_mm512_castps_si512(back), _mm512_castps_si512(x));   // bit patterns differ
      bad |= _mm512_cmpgt_epu32_mask(k, hi);                                // unsigned: negative k included
      b[j] = _mm512_cvtepi32_epi8(k);
    }
    __m512i o = _mm512_castsi128_si512(b[0]);
    o = _mm512_inserti32x4(o, b[1], 1);
    o = _mm512_inserti32x4(o, b[2], 2);
    o = _mm512_inserti32x4(o, b[3], 3);
    if (nt) _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), o);
    else _mm512_storeu_si512(d + i, o);
  }
  if (nt) _mm_sfence();
  bool ok = bad == 0;
  if (i < n) ok &= pack_avx2<NEWTON>(s + i, d + i, n - i);
  return ok;
}

bool pack_range(const float* s, uint8_t* d, int64_t n) {
  static const bool have_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  static const bool have_avx512 = have_avx2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
  static const bool newton = newton_ok();
  if (!have_avx2) return pack_scalar(s, d, n);
  if (have_avx512) return newton ? pack_avx512<true>(s, d, n) : pack_avx512<false>(s, d, n);
  return newton ? pack_avx2<true>(s, d, n) : pack_avx2<false>(s, d, n);
}

// A small persistent pool: thread creation (~30 us each) would otherwise be a visible part of a 1-2 ms job.  Up to MAX_JOBS jobs may be
// in flight: they run one after the other, in submission order, and a worker that finds no chunk left in job j goes straight on to job
// j+1 — the scoring loop submits the pack of batch k+2 while batch k+1 is being packed, so the threads never wait for the caller between
// two packs (measured: a step of the pack-bound pipeline cost the pack time + ~0.5 ms with one job at a time).
class Pool {
 public:
  static constexpr int MAX_JOBS = 2;
  static Pool& get() { static Pool p; return p; }

  // Queues a job for `threads` pool threads and returns; the caller does not take part (it goes on enqueueing GPU work).
  bool begin(const float* src, uint8_t* dst, int64_t n, int threads) {
    std::unique_lock<std::mutex> lk(mu_);
    if (submitted_ - ended_ >= (uint64_t)MAX_JOBS) return false;
    const int64_t items = (n + CHUNK - 1) / CHUNK;
    int want = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    if ((int64_t)want > items) want = (int)items;
    Job& j = jobs_[(submitted_ + 1) % RING];
    j.src = src; j.dst = dst; j.n = n; j.items = items;
    j.next.store(0, std::memory_order_relaxed);
    j.ok.store(true, std::memory_order_relaxed);
    j.started = false;
    j.seq = submitted_ + 1;
    j.want = want;
    j.pending = want;
    j.ms = 0.0;
    while ((int)workers_.size() < want) workers_.emplace_back([this, id = (int)workers_.size(), from = submitted_] { worker(id, from); });
    ++submitted_;
    lk.unlock();
    cv_.notify_all();
    return true;
  }
  // Waits for the OLDEST job in flight; true when every value of it qualified.
  bool end() {
    std::unique_lock<std::mutex> lk(mu_);
    if (submitted_ == ended_) return false;
    Job& j = jobs_[(ended_ + 1) % RING];
    done_cv_.wait(lk, [&] { return j.pending == 0; });
    ++ended_;
    last_ms_ = j.ms;
    return j.ok.load(std::memory_order_relaxed);
  }
  bool idle() { std::unique_lock<std::mutex> lk(mu_); return submitted_ == ended_; }
  double last_ms() { std::unique_lock<std::mutex> lk(mu_); return last_ms_; }

 private:
  static constexpr int RING = 4;       // > MAX_JOBS: a slot is not reused while a worker may still look at it
  struct Job {
    const float* src = nullptr;
    uint8_t* dst = nullptr;
    int64_t n = 0, items = 0;
    std::atomic<int64_t> next{0};
    std::atomic<bool> ok{true};
    bool started = false;              // guarded by mu_
    uint64_t seq = 0;                  // guarded by mu_
    int want = 0, pending = 0;         // guarded by mu_
    std::chrono::steady_clock::time_point t_begin;
    double ms = 0.0;
  };
  Pool() = default;
  ~Pool() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  static void work(Job& j) {
    for (;;) {
      if (!j.ok.load(std::memory_order_relaxed)) return;       // another chunk already failed: stop early
      const int64_t it = j.next.fetch_add(1, std::memory_order_relaxed);
      if (it >= j.items) return;
      const int64_t lo = it * CHUNK, len = (lo + CHUNK <= j.n) ? CHUNK : j.n - lo;
      if (!pack_range(j.src + lo, j.dst + lo, len)) j.ok.store(false, std::memory_order_relaxed);
    }
  }
  void worker(int id, uint64_t seen) {     // seen: sequence number of the last job this worker has dealt with
    for (;;) {
      Job* j = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || submitted_ > seen; });
        if (stop_) return;
        ++seen;
        j = &jobs_[seen % RING];
        // not for this worker: the job runs on fewer threads (a worker that sat one out for so long that the slot was reused —
        // only possible for a job it had no part in, a job cannot end before its participants — skips it as well)
        if (j->seq != seen || id >= j->want) continue;
        if (!j->started) { j->started = true; j->t_begin = std::chrono::steady_clock::now(); }
      }
      work(*j);
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (--j->pending == 0) {
          j->ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - j->t_begin).count();
          done_cv_.notify_all();
        }
      }
    }
  }

  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  uint64_t submitted_ = 0, ended_ = 0;   // sequence numbers: jobs ended_+1 .. submitted_ are in flight
  bool stop_ = false;
  double last_ms_ = 0.0;
  Job jobs_[RING];
};

}  // namespace

extern "C" int lsd_host_pack_u8_exact(const float* src, uint8_t* dst, int64_t n, int threads) {
  if (n < 0 || (n > 0 && (!src || !dst))) return LSD_ERR_ARG;
  if (n == 0) return 1;
  if (!Pool::get().idle() || !Pool::get().begin(src, dst, n, threads)) return LSD_ERR_ARG;     // a begin/end job is in flight
  return Pool::get().end() ? 1 : 0;
}

extern "C" int lsd_host_pack_u8_begin(const float* src, uint8_t* dst, int64_t n, int threads) {
  if (n <= 0 || !src || !dst) return LSD_ERR_ARG;
  return Pool::get().begin(src, dst, n, threads) ? LSD_OK : LSD_ERR_ARG;
}

extern "C" int lsd_host_pack_u8_end(void) { return Pool::get().end() ? 1 : 0; }

extern "C" double lsd_host_pack_last_ms(void) { return Pool::get().last_ms(); }
