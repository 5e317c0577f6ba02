// Host-side bit packing of soft masks (packed mask formats).
//
// The IoU part of the cost needs only the thresholded bits (match_helper.py:16-17: `mask > 0.5`).  When the masks
// live in HOST memory (the e2e / plugin-with-host-buffers case), shipping fp32 over PCIe costs 27.5 MB per match;
// packing them here -- a streaming compare the host cores do at memory speed -- ships 0.86 MB instead.
// bit i of word j of a row = (pixel 32*j + i) > 0.5f; bits past the row end are 0.  NaN compares false, like torch.
#include <stdint.h>
#include <stddef.h>

#include <atomic>
#include <functional>
#include <thread>

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
  // Plain std::threads per call, work claimed in blocks from an atomic counter.  (Round 1 used an OpenMP team: its idle
  // workers SPIN after the parallel region, which under a cgroup CPU quota -- the GPU boxes give 128 logical CPUs a
  // 16-CPU quota -- burns the quota and gets the whole process, kernel-launching thread included, throttled.)
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
  // The team starts as a binary tree (thread i starts 2i+1 and 2i+2 before it works): the last of 16 threads is running
  // after 4 thread creations instead of 15, which matters when the whole call is ~10 ms.
  std::function<void(long long)> run = [&](long long id) {
    std::thread c1, c2;
    if (2 * id + 1 < nt) c1 = std::thread(run, 2 * id + 1);
    if (2 * id + 2 < nt) c2 = std::thread(run, 2 * id + 2);
    work();
    if (c1.joinable()) c1.join();
    if (c2.joinable()) c2.join();
  };
  run(0);
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
