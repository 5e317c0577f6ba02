/* Plain-C restatement of the integer part of the mask-IoU cost.  TEST INFRASTRUCTURE ONLY (see oracle/match_oracle.py).
 *
 * Reference: dmm/utils/match_helper.py:9-28 (compute_iou_binary_mask_2D: rows thresholded `> 0.5`, intersection and union
 * as sums of the and / or of the binary rows, iou = inter / (union + 1e-6) in fp32) applied to every (template, proposal)
 * pair as dmm/modules/match_model.py:83-89 expands them.  No SIMD, no bit tricks: one compare per pixel, integer counts.
 * Only tests/ (and build(), which compiles it) touch this file; it cross-checks the torch-CPU oracle and the CUDA kernels
 * with an independent implementation of the same integers.
 *
 *   gcc -O2 -shared -fPIC oracle/iou_oracle.c -o oracle/_build/libiou_oracle.so       (done by __graft_entry__.build())
 */
#include <stdint.h>

/* prop [P][HW], tmpl [O][HW] fp32 -> inter [O][P], area_t [O], area_p [P] (int64), iou [O][P] (fp32) */
int iou_oracle_pairwise(const float* prop, const float* tmpl, int P, int O, long long HW, int64_t* inter, int64_t* area_t,
                        int64_t* area_p, float* iou) {
  if (P < 0 || O < 0 || HW < 0) return 1;
  for (int p = 0; p < P; ++p) {
    int64_t a = 0;
    for (long long i = 0; i < HW; ++i) a += prop[(long long)p * HW + i] > 0.5f;
    area_p[p] = a;
  }
  for (int o = 0; o < O; ++o) {
    int64_t a = 0;
    for (long long i = 0; i < HW; ++i) a += tmpl[(long long)o * HW + i] > 0.5f;
    area_t[o] = a;
    for (int p = 0; p < P; ++p) {
      int64_t n = 0;
      for (long long i = 0; i < HW; ++i) n += (tmpl[(long long)o * HW + i] > 0.5f) & (prop[(long long)p * HW + i] > 0.5f);
      inter[(long long)o * P + p] = n;
      /* union = |A| + |B| - |A & B| (an exact integer below 2^24 in the reference's fp32 sum), then the two fp32 operations */
      const float uni = (float)(area_t[o] + area_p[p] - n) + 1e-6f;
      iou[(long long)o * P + p] = (float)n / uni;
    }
  }
  return 0;
}

/* the packed-row format of the library: bit i of word j = pixel 32*j + i > 0.5, zero tail */
int iou_oracle_pack_bits(const float* rows, long long n_rows, long long HW, uint32_t* bits) {
  const long long words = (HW + 31) / 32;
  for (long long r = 0; r < n_rows; ++r)
    for (long long j = 0; j < words; ++j) {
      uint32_t w = 0;
      for (int i = 0; i < 32; ++i) {
        const long long px = 32 * j + i;
        if (px < HW && rows[r * HW + px] > 0.5f) w |= 1u << i;
      }
      bits[r * words + j] = w;
    }
  return 0;
}
