// Host-side bit packing of soft masks (packed mask formats).
//
// The IoU part of the cost needs only the thresholded bits (match_helper.py:16-17: `mask > 0.5`).  When the masks
// live in HOST memory (the e2e / plugin-with-host-buffers case), shipping fp32 over PCIe costs 27.5 MB per match;
// packing them here -- a streaming compare the host cores do at memory speed -- ships 0.86 MB instead.
// bit i of word j of a row = (pixel 32*j + i) > 0.5f; bits past the row end are 0.  NaN compares false, like torch.
#include <stdint.h>
#include <stddef.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/dmm_b200.h"

namespace {

inline uint32_t pack32_scalar(const float* x, long long n) {
  uint32_t w = 0;
  for (long long i = 0; i < n; ++i) w |= (uint32_t)(x[i] > 0.5f) << i;
  return w;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_row_avx2(const float* src, long long HW, uint32_t* dst, long long words) {
  const __m256 half = _mm256_set1_ps(0.5f);
  const long long full = HW / 32;
  for (long long j = 0; j < full; ++j) {
    const float* x = src + 32 * j;
    const uint32_t b0 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(x), half, _CMP_GT_OQ));
    const uint32_t b1 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(x + 8), half, _CMP_GT_OQ));
    const uint32_t b2 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(x + 16), half, _CMP_GT_OQ));
    const uint32_t b3 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(x + 24), half, _CMP_GT_OQ));
    dst[j] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
  }
  if (full < words) dst[full] = pack32_scalar(src + 32 * full, HW - 32 * full);
}
#endif

#if defined(__x86_64__)
__attribute__((target("avx512f"))) void pack_row_avx512(const float* src, long long HW, uint32_t* dst, long long words) {
  const __m512 half = _mm512_set1_ps(0.5f);
  const long long full = HW / 32;
  for (long long j = 0; j < full; ++j) {
    const float* x = src + 32 * j;
    const uint32_t b0 = (uint32_t)_mm512_cmp_ps_mask(_mm512_loadu_ps(x), half, _CMP_GT_OQ);
    const uint32_t b1 = (uint32_t)_mm512_cmp_ps_mask(_mm512_loadu_ps(x + 16), half, _CMP_GT_OQ);
    dst[j] = b0 | (b1 << 16);
  }
  if (full < words) dst[full] = pack32_scalar(src + 32 * full, HW - 32 * full);
}
#endif

void pack_row_scalar(const float* src, long long HW, uint32_t* dst, long long words) {
  const long long full = HW / 32;
  for (long long j = 0; j < full; ++j) dst[j] = pack32_scalar(src + 32 * j, 32);
  if (full < words) dst[full] = pack32_scalar(src + 32 * full, HW - 32 * full);
}

}  // namespace

extern "C" long long dmm_packed_words(long long HW) { return (HW + 31) / 32; }

namespace {

// Persistent worker team: threads are created on first use, PARK on a condition variable between calls (no spinning: see
// pack_jobs) and live until the process exits.  run(n, fn) executes fn on n threads (the caller is one of them) and
// returns when all have finished.  One call at a time per process (a mutex serialises concurrent callers).
class Team {
 public:
  void run(int n, const std::function<void()>& fn) {
    std::lock_guard<std::mutex> call(call_mu_);
    if (n <= 1) { fn(); return; }
    {
      std::unique_lock<std::mutex> lk(mu_);
      while ((int)workers_.size() < n - 1) workers_.emplace_back([this, id = (int)workers_.size()] { loop(id); });
      fn_ = &fn; want_ = n - 1; pending_ = n - 1; ++gen_;
    }
    cv_.notify_all();
    fn();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }
  ~Team() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) if (t.joinable()) t.join();
  }

 private:
  void loop(int id) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void()>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        if (id >= want_) continue;                           // this call uses fewer threads
        fn = fn_;
      }
      (*fn)();
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> workers_;
  const std::function<void()>* fn_ = nullptr;
  unsigned long long gen_ = 0;
  int want_ = 0, pending_ = 0;
  bool stop_ = false;
};

Team& team() {
  static Team t;
  return t;
}

struct PackJob { const float* src; long long rows; uint32_t* dst; };

int pack_jobs(const PackJob* jobs, int njobs, long long HW, int threads) {
  const long long words = (HW + 31) / 32;
#if defined(__x86_64__)
  const bool avx2 = __builtin_cpu_supports("avx2");
  const bool avx512 = __builtin_cpu_supports("avx512f");
#else
  const bool avx2 = false, avx512 = false;
#endif
  // split every row into pieces so that a handful of huge rows still spreads over all threads
  const long long piece_words = 1024;  // 128 KB of fp32 per task
  const long long pieces = (words + piece_words - 1) / piece_words;
  long long first[3] = {0, 0, 0};      // task ranges of the (at most two) jobs
  for (int j = 0; j < njobs; ++j) first[j + 1] = first[j] + jobs[j].rows * pieces;
  const long long tasks = first[njobs];
  if (tasks == 0) return DMM_OK;
  // A parked team (Team below) runs the blocks; work is claimed from an atomic counter.  History: round 1 used an OpenMP
  // team, whose idle workers SPIN after the parallel region -- under the GPU boxes' cgroup CPU quota (128 logical CPUs,
  // 16-CPU quota) that burned the quota and got the whole process throttled; then plain std::threads created per call
  // (tree start-up), which put 16 thread creations into every ~10 ms step.  Now the workers are created once and sleep
  // on a condition variable between calls.
  long long nt = threads > 0 ? threads : (long long)std::thread::hardware_concurrency();
  const long long block = 8;           // tasks per claim: 1 MB of fp32
  if (nt > (tasks + block - 1) / block) nt = (tasks + block - 1) / block;
  if (nt < 1) nt = 1;
  std::atomic<long long> next(0);
  auto work = [&]() {
    for (;;) {
      const long long t0 = next.fetch_add(block, std::memory_order_relaxed);
      if (t0 >= tasks) break;
      const long long t1 = t0 + block < tasks ? t0 + block : tasks;
      for (long long tt = t0; tt < t1; ++tt) {
        const int j = (njobs > 1 && tt >= first[1]) ? 1 : 0;
        const long long t = tt - first[j];
        const long long r = t / pieces, pc = t - r * pieces;
        const long long w0 = pc * piece_words;
        const long long w1 = w0 + piece_words < words ? w0 + piece_words : words;
        const float* s = jobs[j].src + r * HW + 32 * w0;
        const long long n = (32 * w1 < HW ? 32 * w1 : HW) - 32 * w0;
        uint32_t* d = jobs[j].dst + r * words + w0;
#if defined(__x86_64__)
        if (avx512) { pack_row_avx512(s, n, d, w1 - w0); continue; }
        if (avx2) { pack_row_avx2(s, n, d, w1 - w0); continue; }
#endif
        pack_row_scalar(s, n, d, w1 - w0);
      }
    }
  };
  team().run((int)nt, work);
  return DMM_OK;
}

}  // namespace

extern "C" int dmm_host_pack_masks(const float* src, long long rows, long long HW, uint32_t* dst, int threads) {
  if (rows < 0 || HW < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (rows == 0 || HW == 0) return DMM_OK;
  if (!src || !dst) return DMM_ERR_INVALID_ARGUMENT;
  const PackJob job = {src, rows, dst};
  return pack_jobs(&job, 1, HW, threads);
}

extern "C" int dmm_host_pack_masks2(const float* src_a, long long rows_a, uint32_t* dst_a, const float* src_b,
                                    long long rows_b, uint32_t* dst_b, long long HW, int threads) {
  if (rows_a < 0 || rows_b < 0 || HW < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (HW == 0) return DMM_OK;
  if ((rows_a > 0 && (!src_a || !dst_a)) || (rows_b > 0 && (!src_b || !dst_b))) return DMM_ERR_INVALID_ARGUMENT;
  const PackJob jobs[2] = {{src_a, rows_a, dst_a}, {src_b, rows_b, dst_b}};
  return pack_jobs(jobs, 2, HW, threads);
}

// Streaming-read bandwidth of a host buffer with the packer's thread team (bench.py: the host-DRAM roofline of the
// host-buffer entry -- every mask byte has to be read from host memory once, whichever route carries it).  Each thread
// sums 64-byte lines of its blocks; returns GB/s of the best of `reps` passes through *gbs.
extern "C" int dmm_host_read_bandwidth(const void* src, long long bytes, int threads, int reps, double* gbs) {
  if (!src || !gbs || bytes <= 0 || reps <= 0) return DMM_ERR_INVALID_ARGUMENT;
  const long long block = 1 << 20, nblocks = (bytes + block - 1) / block;
  long long nt = threads > 0 ? threads : (long long)std::thread::hardware_concurrency();
  if (nt > nblocks) nt = nblocks;
  double best = 0.0;
  std::atomic<unsigned long long> sink(0);
  for (int r = 0; r < reps; ++r) {
    std::atomic<long long> next(0);
    auto work = [&]() {
      unsigned long long acc = 0;
      for (;;) {
        const long long b = next.fetch_add(1, std::memory_order_relaxed);
        if (b >= nblocks) break;
        const long long lo = b * block, hi = lo + block < bytes ? lo + block : bytes;
        const unsigned long long* q = (const unsigned long long*)((const char*)src + lo);
        const long long n = (hi - lo) / 8;
        unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        long long i = 0;
        for (; i + 32 <= n; i += 32) {                       // 4 cache lines per iteration, every word touched
          for (int k = 0; k < 8; ++k) { a0 += q[i + k]; a1 += q[i + 8 + k]; a2 += q[i + 16 + k]; a3 += q[i + 24 + k]; }
        }
        for (; i < n; ++i) a0 += q[i];
        acc += a0 + a1 + a2 + a3;
      }
      sink.fetch_add(acc, std::memory_order_relaxed);
    };
    const auto t0 = std::chrono::steady_clock::now();
    team().run((int)nt, work);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (sec > 0 && bytes / sec / 1e9 > best) best = bytes / sec / 1e9;
  }
  *gbs = best + (sink.load() == 0xdeadbeefULL ? 1e-30 : 0.0);   // keep the sums alive
  return DMM_OK;
}
