// Host-side bit packing of soft masks (packed mask formats).
//
// The IoU part of the cost needs only the thresholded bits (match_helper.py:16-17: `mask > 0.5`).  When the masks
// live in HOST memory (the e2e / plugin-with-host-buffers case), shipping fp32 over PCIe costs 27.5 MB per match;
// packing them here -- a streaming compare the host cores do at memory speed -- ships 0.86 MB instead.
// bit i of word j of a row = (pixel 32*j + i) > 0.5f; bits past the row end are 0.  NaN compares false, like torch.
#include <stdint.h>
#include <stddef.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
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

void pack_row_scalar(const float* src, long long HW, uint32_t* dst, long long words) {
  const long long full = HW / 32;
  for (long long j = 0; j < full; ++j) dst[j] = pack32_scalar(src + 32 * j, 32);
  if (full < words) dst[full] = pack32_scalar(src + 32 * full, HW - 32 * full);
}

}  // namespace

extern "C" long long dmm_packed_words(long long HW) { return (HW + 31) / 32; }

extern "C" int dmm_host_pack_masks(const float* src, long long rows, long long HW, uint32_t* dst, int threads) {
  if (rows < 0 || HW < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (rows == 0 || HW == 0) return DMM_OK;
  if (!src || !dst) return DMM_ERR_INVALID_ARGUMENT;
  const long long words = (HW + 31) / 32;
#if defined(__x86_64__)
  const bool avx2 = __builtin_cpu_supports("avx2");
#else
  const bool avx2 = false;
#endif
  // split every row into pieces so that a handful of huge rows still spreads over all threads
  const long long piece_words = 1024;  // 128 KB of fp32 per task
  const long long pieces = (words + piece_words - 1) / piece_words;
  const long long tasks = rows * pieces;
#ifdef _OPENMP
  const int nt = threads > 0 ? threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
  for (long long t = 0; t < tasks; ++t) {
    const long long r = t / pieces, pc = t - r * pieces;
    const long long w0 = pc * piece_words;
    const long long w1 = w0 + piece_words < words ? w0 + piece_words : words;
    const float* s = src + r * HW + 32 * w0;
    const long long n = (32 * w1 < HW ? 32 * w1 : HW) - 32 * w0;
    uint32_t* d = dst + r * words + w0;
#if defined(__x86_64__)
    if (avx2) { pack_row_avx2(s, n, d, w1 - w0); continue; }
#endif
    pack_row_scalar(s, n, d, w1 - w0);
  }
  (void)threads;
  return DMM_OK;
}
