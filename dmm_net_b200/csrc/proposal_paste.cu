// K8 -- proposal mask paste (+ bit planes + tight boxes), K9 -- box NMS.   SURVEY.md section 8f-4.
//
// Reference K8: dmm/utils/masker.py:91-173 (expand_boxes, expand_masks, paste_mask_in_image, binmask_to_box) looped over
// the proposals of an image by Masker.forward_single_image (masker.py:181-206): per proposal a Python-side sequence of
// zero-pad, int box, F.interpolate(bilinear, align_corners=False) to the box size, slice-assign into a zero image,
// nonzero() + 4 .item() syncs for the tight box.  Here ONE launch writes every pasted soft mask [N][im_h][im_w] of a
// batch of frames -- exactly the dense fp32 rows K1/K4 read -- and, from the same registers, the thresholded bit rows of
// the packed K1 entry and the tight boxes.  HBM-write-bound: N*im_h*im_w*4 bytes written once (most of them zeros
// outside the box), the M x M source mask sits in shared memory.
// Arithmetic follows ATen's upsample_bilinear2d (source index = scale*(dst+0.5)-0.5 clamped at 0, lambdas, the
// four-tap blend w0*a + w1*b) with the contraction ATen's x86 build applies -- fma(scale, dst+0.5, -0.5) and
// fma(w0, a, w1*b).  Measured against the golden vectors of the CPU reference: 99.1 % of the pixels bit-equal, the
// rest within one ulp (ATen has several bilinear loops, chosen by strides and sizes, that its compiler contracts
// differently; without the fmas 22 % of the pixels differ in the last bit and pixels whose source index is 0 up to
// rounding flip between exact 0 and 1e-9).  Tests allow 1e-6.
//
// Reference K9: dmm/utils/boxlist_ops.py:15-29 -> maskrcnn_benchmark.layers.nms (un-vendored; parity unpinned): greedy
// score-descending suppression with the legacy +1 pixel widths.
#include <limits.h>

#include "common.cuh"

namespace dmm {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxMp = 64;  // padded source mask side (M + 2*padding)

struct PasteParams {
  const float* masks;    // [N][M][M]
  const float* boxes;    // [N][4] xyxy
  float* out;            // [N][im_h][im_w] or NULL
  uint32_t* bits;        // [N][words] (zero-initialised by the entry point) or NULL
  int* tight_ws;         // [N][4] = xmin, ymin, -xmax, -ymax (atomicMin, sentinel 0x7f7f7f7f) or NULL
  int N, M, pad, im_h, im_w, words, rows_per_cta;
  float scale, thresh;
};

struct BoxI { int x0, y0, x1, y1, w, h; };

// masker.py:91-108 + :124: expand around the centre in fp32, truncate to int32
__device__ __forceinline__ BoxI expand_box(const float* b, float scale) {
  const float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
  const float w_half = __fmul_rn(__fmul_rn(__fsub_rn(x2, x1), 0.5f), scale);
  const float h_half = __fmul_rn(__fmul_rn(__fsub_rn(y2, y1), 0.5f), scale);
  const float x_c = __fmul_rn(__fadd_rn(x2, x1), 0.5f);
  const float y_c = __fmul_rn(__fadd_rn(y2, y1), 0.5f);
  BoxI r;
  r.x0 = (int)__fsub_rn(x_c, w_half); r.x1 = (int)__fadd_rn(x_c, w_half);
  r.y0 = (int)__fsub_rn(y_c, h_half); r.y1 = (int)__fadd_rn(y_c, h_half);
  r.w = max(r.x1 - r.x0 + 1, 1);
  r.h = max(r.y1 - r.y0 + 1, 1);
  return r;
}

struct Tap { int i0, i1; float l0, l1; };
// ATen area_pixel_compute_source_index (align_corners=False) + the index / lambda split
__device__ __forceinline__ Tap tap(int dst, float scale, int in_size) {
  float s = fmaf(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
  s = s < 0.f ? 0.f : s;
  Tap t;
  t.i0 = min((int)s, in_size - 1);
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = __fsub_rn(s, (float)t.i0);
  t.l0 = __fsub_rn(1.f, t.l1);
  return t;
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads) paste_masks_kernel(const PasteParams p) {
  __shared__ float sm[kMaxMp * kMaxMp];
  __shared__ int s_box[4];
  const int n = blockIdx.y, tid = threadIdx.x;
  const int Mp = p.M + 2 * p.pad;
  const BoxI bx = expand_box(p.boxes + 4LL * n, p.scale);
  const int row_lo = blockIdx.x * p.rows_per_cta, row_hi = min(row_lo + p.rows_per_cta, p.im_h);
  const int x0c = max(bx.x0, 0), x1c = min(bx.x1 + 1, p.im_w);
  const int y0c = max(bx.y0, 0), y1c = min(bx.y1 + 1, p.im_h);
  const bool cta_touches = y0c < row_hi && y1c > row_lo && x0c < x1c;
  if (cta_touches) {   // block-uniform
    const float* src = p.masks + (long long)n * p.M * p.M;
    for (int i = tid; i < Mp * Mp; i += kThreads) {
      const int y = i / Mp - p.pad, x = i % Mp - p.pad;
      sm[i] = (y >= 0 && y < p.M && x >= 0 && x < p.M) ? src[y * p.M + x] : 0.f;   // expand_masks: zero border
    }
    if (tid < 4) s_box[tid] = INT_MAX;
    __syncthreads();
  }
  const float sc_y = (float)Mp / (float)bx.h, sc_x = (float)Mp / (float)bx.w;      // area_pixel_compute_scale
  constexpr int step = VEC ? 4 : 1;
  float* outn = p.out ? p.out + (long long)n * p.im_h * p.im_w : nullptr;
  // with soft rows to write, a CTA walks its whole slab (zeros outside the box); for bit rows / tight boxes alone only
  // the part of the box inside the slab matters (the lazy pipeline: ~10x fewer items)
  const int ya = outn ? row_lo : max(row_lo, y0c), yb = outn ? row_hi : min(row_hi, y1c);
  const int g0 = outn ? 0 : x0c / step, g1 = outn ? (p.im_w + step - 1) / step : (x1c + step - 1) / step;
  const int groups = max(g1 - g0, 0);
  const int items = cta_touches || outn ? max(yb - ya, 0) * groups : 0;
  int t_xmin = INT_MAX, t_ymin = INT_MAX, t_nxmax = INT_MAX, t_nymax = INT_MAX;
  // item -> (row, column group) without an integer division per item: both kernels here are ISSUE-bound (ncu: 85 % issue
  // slots busy, DRAM 63 %), and the two divisions were most of what a zero-fill item costs.  __umulhi(it, inv) == it / groups
  // exactly while it * groups < 2^32 (always, short of ~4k-wide images: then the plain division is used).
  const unsigned inv = groups > 0 ? 0xFFFFFFFFu / (unsigned)groups + 1u : 0u;
  const bool fast_div = (unsigned long long)items * (unsigned)groups < 0xFFFFFFFFull && groups > 1;
  for (int it = tid; it < items; it += kThreads) {
    const int row = fast_div ? (int)__umulhi((unsigned)it, inv) : (groups > 0 ? it / groups : 0);
    const int Y = ya + row, X = (g0 + it - row * groups) * step;
    float v[step];
#pragma unroll
    for (int k = 0; k < step; ++k) v[k] = 0.f;
    if (cta_touches && Y >= y0c && Y < y1c && X + step > x0c && X < x1c) {
      const Tap ty = tap(Y - bx.y0, sc_y, Mp);
      const float* r0 = sm + ty.i0 * Mp;
      const float* r1 = sm + ty.i1 * Mp;
#pragma unroll
      for (int k = 0; k < step; ++k) {
        const int Xk = X + k;
        if (Xk >= x0c && Xk < x1c) {
          const Tap tx = tap(Xk - bx.x0, sc_x, Mp);
          const float top = fmaf(tx.l0, r0[tx.i0], __fmul_rn(tx.l1, r0[tx.i1]));
          const float bot = fmaf(tx.l0, r1[tx.i0], __fmul_rn(tx.l1, r1[tx.i1]));
          v[k] = fmaf(ty.l0, top, __fmul_rn(ty.l1, bot));
          if (v[k] > p.thresh) {
            t_xmin = min(t_xmin, Xk); t_nxmax = min(t_nxmax, -Xk);
            t_ymin = min(t_ymin, Y); t_nymax = min(t_nymax, -Y);
          }
        }
      }
      if (p.bits) {
        const long long px = (long long)Y * p.im_w + X;        // VEC: im_w % 4 == 0, so a nibble never straddles a word
        unsigned nib = 0;
#pragma unroll
        for (int k = 0; k < step; ++k) nib |= (v[k] > 0.5f ? 1u : 0u) << k;
        if (nib) atomicOr(p.bits + (long long)n * p.words + (px >> 5), nib << (px & 31));
      }
    }
    if (outn) {
      if constexpr (VEC) st_stream_f4(outn + (long long)Y * p.im_w + X, make_float4(v[0], v[1], v[2], v[3]));
      else outn[(long long)Y * p.im_w + X] = v[0];
    }
  }
  if (cta_touches && p.tight_ws) {
    // warp min, then one shared atomic per warp, one global atomic per CTA and coordinate
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t_xmin = min(t_xmin, __shfl_xor_sync(0xffffffffu, t_xmin, o));
      t_ymin = min(t_ymin, __shfl_xor_sync(0xffffffffu, t_ymin, o));
      t_nxmax = min(t_nxmax, __shfl_xor_sync(0xffffffffu, t_nxmax, o));
      t_nymax = min(t_nymax, __shfl_xor_sync(0xffffffffu, t_nymax, o));
    }
    if ((tid & 31) == 0 && t_xmin != INT_MAX) {
      atomicMin(&s_box[0], t_xmin); atomicMin(&s_box[1], t_ymin);
      atomicMin(&s_box[2], t_nxmax); atomicMin(&s_box[3], t_nymax);
    }
    __syncthreads();
    if (tid < 4 && s_box[tid] != INT_MAX) atomicMin(p.tight_ws + 4LL * n + tid, s_box[tid]);
  }
}

// binmask_to_box (masker.py:152-166): [xmin, ymin, xmax, ymax], or [0, 0, im_h, im_w] for an empty mask (sic)
__global__ void paste_tight_finalize_kernel(const int* __restrict__ ws, long long* __restrict__ tight, int N, int im_h,
                                            int im_w) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int4 w = *reinterpret_cast<const int4*>(ws + 4LL * n);
  long long* t = tight + 4LL * n;
  if (w.x == 0x7f7f7f7f) { t[0] = 0; t[1] = 0; t[2] = im_h; t[3] = im_w; }
  else { t[0] = w.x; t[1] = w.y; t[2] = -w.z; t[3] = -w.w; }
}

// ---- K10: fused paste + assignment apply ("lazy paste") ---------------------------------------------------------
// out[b, row(o), :] = sum_p Bmat[b,o,p] * paste(mask[src[b,p]], box[src[b,p]])      (match_model.py:144 on top of masker.py)
// The eval pipeline pastes P soft masks per frame (P*HW*4 bytes written), reads them all back for the IoU and once more
// for the apply, although the IoU needs one bit per pixel (K8's bit rows -> packed K1) and the apply needs the <= O
// SELECTED proposals.  This kernel pastes only those, scaled, straight into the output rows: per frame O*HW*4 bytes
// written and nothing read but the 28x28 heads.  Same tap arithmetic as K8 and the same fmaf(v, m, acc) accumulation in
// ascending proposal order as K4, so the result is bit-identical to K8 followed by K4.
constexpr int kMaxSrc = 8;       // source masks resident in shared memory per pass

struct PasteApplyParams {
  const float* coef;             // [B][O][MS]
  const float* masks;            // [Nsrc][M][M]
  const float* boxes;            // [Nsrc][4]
  const int* src;                // [B][P] row of masks/boxes for column p, or < 0
  const int* n_prop; const int* n_tmpl; const int* row_map;
  float* out; long long out_bs;
  int B, P, O, MS, O_out, M, pad, im_h, im_w, rows_per_cta, zero_fill;
  float scale;
};

template <bool VEC>
__global__ void __launch_bounds__(kThreads) paste_apply_kernel(const PasteApplyParams p) {
  extern __shared__ float sm_src[];               // [kMaxSrc][Mp*Mp]
  __shared__ int s_col[128];                      // non-zero columns of this output row, ascending
  __shared__ float s_val[128];
  __shared__ int s_cnt, s_o;
  __shared__ BoxI s_bx[kMaxSrc];
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.y / p.O_out, f = blockIdx.y % p.O_out;
  const int Mp = p.M + 2 * p.pad;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  if (tid == 0) {                                  // which template row lands in output row f (row_map may scatter)
    int o = -1;
    if (!p.row_map) o = f < nt ? f : -1;
    else for (int k = 0; k < nt; ++k) if (p.row_map[(long long)b * p.O + k] == f) o = k;
    s_o = o;
    s_cnt = 0;
  }
  __syncthreads();
  const int o = s_o;
  if (o < 0 && !p.zero_fill) return;
  if (o >= 0 && tid < 32) {                        // warp 0: ballot-compact the non-zero coefficients of row o
    const float* crow = p.coef + ((long long)b * p.O + o) * p.MS;
    int cnt = 0;
    for (int c0 = 0; c0 < np; c0 += 32) {
      const int c = c0 + lane;
      const float v = c < np ? crow[c] : 0.f;
      const bool nz = v != 0.f && p.src[(long long)b * p.P + min(c, p.P - 1)] >= 0 && c < np;
      const unsigned mk = __ballot_sync(0xffffffffu, nz);
      if (nz) { const int pos = cnt + __popc(mk & ((1u << lane) - 1u)); s_col[pos] = c; s_val[pos] = v; }
      cnt += __popc(mk);
    }
    if (lane == 0) s_cnt = cnt;
  }
  __syncthreads();
  const int cnt = s_cnt;
  const int row_lo = blockIdx.x * p.rows_per_cta, row_hi = min(row_lo + p.rows_per_cta, p.im_h);
  constexpr int step = VEC ? 4 : 1;
  const int groups = (p.im_w + step - 1) / step;
  const int items = (row_hi - row_lo) * groups;
  const unsigned inv = 0xFFFFFFFFu / (unsigned)groups + 1u;
  const bool fast_div = (unsigned long long)items * (unsigned)groups < 0xFFFFFFFFull && groups > 1;
  float* orow = p.out + (long long)b * p.out_bs + (long long)f * p.im_h * p.im_w;
  for (int e0 = 0; e0 == 0 || e0 < cnt; e0 += kMaxSrc) {     // one pass per kMaxSrc sources (a single pass in eval mode)
    const int ne = min(kMaxSrc, cnt - e0);
    __syncthreads();
    for (int k = 0; k < ne; ++k) {
      const int srow = p.src[(long long)b * p.P + s_col[e0 + k]];
      const float* src = p.masks + (long long)srow * p.M * p.M;
      for (int i = tid; i < Mp * Mp; i += kThreads) {
        const int y = i / Mp - p.pad, x = i % Mp - p.pad;
        sm_src[k * Mp * Mp + i] = (y >= 0 && y < p.M && x >= 0 && x < p.M) ? src[y * p.M + x] : 0.f;
      }
      if (tid == 0) s_bx[k] = expand_box(p.boxes + 4LL * srow, p.scale);
    }
    __syncthreads();
    for (int it = tid; it < items; it += kThreads) {
      const int row = fast_div ? (int)__umulhi((unsigned)it, inv) : it / groups;      // see paste_masks_kernel
      const int Y = row_lo + row, X = (it - row * groups) * step;
      float acc[step];
      if (e0 == 0) {
#pragma unroll
        for (int k = 0; k < step; ++k) acc[k] = 0.f;
      } else {                                      // later passes accumulate on top of what the earlier ones wrote
        if constexpr (VEC) {
          const float4 q = *reinterpret_cast<const float4*>(orow + (long long)Y * p.im_w + X);
          acc[0] = q.x; acc[1] = q.y; acc[2] = q.z; acc[3] = q.w;
        } else {
          acc[0] = orow[(long long)Y * p.im_w + X];
        }
      }
      for (int k = 0; k < ne; ++k) {
        const BoxI bx = s_bx[k];
        const int x0c = max(bx.x0, 0), x1c = min(bx.x1 + 1, p.im_w);
        const int y0c = max(bx.y0, 0), y1c = min(bx.y1 + 1, p.im_h);
        if (Y < y0c || Y >= y1c || X + step <= x0c || X >= x1c) continue;
        const float v = s_val[e0 + k];
        const float sc_y = (float)Mp / (float)bx.h, sc_x = (float)Mp / (float)bx.w;
        const Tap ty = tap(Y - bx.y0, sc_y, Mp);
        const float* r0 = sm_src + k * Mp * Mp + ty.i0 * Mp;
        const float* r1 = sm_src + k * Mp * Mp + ty.i1 * Mp;
#pragma unroll
        for (int j = 0; j < step; ++j) {
          const int Xj = X + j;
          if (Xj >= x0c && Xj < x1c) {
            const Tap tx = tap(Xj - bx.x0, sc_x, Mp);
            const float top = fmaf(tx.l0, r0[tx.i0], __fmul_rn(tx.l1, r0[tx.i1]));
            const float bot = fmaf(tx.l0, r1[tx.i0], __fmul_rn(tx.l1, r1[tx.i1]));
            const float m = fmaf(ty.l0, top, __fmul_rn(ty.l1, bot));
            acc[j] = cnt == 1 ? __fmul_rn(v, m) : fmaf(v, m, acc[j]);      // K4: scaled copy for one source, fmaf chain else
          }
        }
      }
      if constexpr (VEC) st_stream_f4(orow + (long long)Y * p.im_w + X, make_float4(acc[0], acc[1], acc[2], acc[3]));
      else orow[(long long)Y * p.im_w + X] = acc[0];
    }
  }
}

// ---- K9 ------------------------------------------------------------------------------------------------------
constexpr int kNmsMax = 1024;

__global__ void __launch_bounds__(kNmsMax) box_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                          const int* __restrict__ n_boxes, int n_max, float thresh,
                                                          int max_keep, long long* __restrict__ keep,
                                                          int* __restrict__ n_keep) {
  __shared__ float s_key[kNmsMax];
  __shared__ int s_idx[kNmsMax];
  __shared__ float4 s_box[kNmsMax];
  __shared__ unsigned char s_removed[kNmsMax];
  __shared__ int s_count;
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = n_boxes ? clampi(n_boxes[f], 0, n_max) : n_max;
  const float* bf = boxes + (long long)f * n_max * 4;
  const float* sf = scores + (long long)f * n_max;
  // bitonic sort of (score desc, index asc); padding sorts last
  s_key[tid] = tid < n ? sf[tid] : -INFINITY;
  s_idx[tid] = tid < n ? tid : INT_MAX;
  __syncthreads();
  for (int k = 2; k <= kNmsMax; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int other = tid ^ j;
      if (other > tid) {
        const float a = s_key[tid], b = s_key[other];
        const int ia = s_idx[tid], ib = s_idx[other];
        const bool a_first = a > b || (a == b && ia < ib) || (b != b && a == a);   // does a sort before b?
        const bool up = (tid & k) == 0;
        if (up ? !a_first : a_first) {
          s_key[tid] = b; s_key[other] = a; s_idx[tid] = ib; s_idx[other] = ia;
        }
      }
      __syncthreads();
    }
  }
  if (tid < n) s_box[tid] = *reinterpret_cast<const float4*>(bf + 4LL * s_idx[tid]);
  s_removed[tid] = tid < n ? 0 : 1;
  if (tid == 0) s_count = 0;
  __syncthreads();
  float4 me = tid < n ? s_box[tid] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float my_area = __fmul_rn(__fadd_rn(__fsub_rn(me.z, me.x), 1.f), __fadd_rn(__fsub_rn(me.w, me.y), 1.f));
  long long* kf = keep + (long long)f * n_max;
  for (int i = 0; i < n; ++i) {
    if (s_removed[i]) continue;           // block-uniform (shared, read after the barrier below)
    if (tid == 0) {
      if (max_keep <= 0 || s_count < max_keep) kf[s_count] = s_idx[i];
      s_count++;
    }
    if (tid > i && tid < n && !s_removed[tid]) {
      const float4 a = s_box[i];
      const float a_area = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
      const float left = fmaxf(a.x, me.x), right = fminf(a.z, me.z);
      const float top = fmaxf(a.y, me.y), bottom = fminf(a.w, me.w);
      const float iw = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
      const float ih = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
      const float inter = __fmul_rn(iw, ih);
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a_area, my_area), inter));
      if (iou > thresh) s_removed[tid] = 1;
    }
    __syncthreads();
  }
  __syncthreads();
  const int kept = (max_keep > 0 && max_keep < n) ? min(s_count, max_keep) : s_count;
  if (tid == 0) n_keep[f] = kept;
  for (int i = kept + tid; i < n_max; i += blockDim.x) kf[i] = -1;
}

inline bool aligned16(const void* q) { return ((uintptr_t)q & 15u) == 0; }

}  // namespace
}  // namespace dmm

using namespace dmm;

extern "C" size_t dmm_paste_masks_workspace_bytes(int N) { return align_up((size_t)(N > 0 ? N : 0) * 4 * sizeof(int), 256); }

extern "C" int dmm_paste_masks(const float* masks, const float* boxes, int N, int M, int padding, int im_h, int im_w,
                               float thresh, float* pasted, uint32_t* bits, long long* tight, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (N < 0 || M <= 0 || padding < 1 || im_h < 0 || im_w < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (M + 2 * padding > kMaxMp || N > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;
  if (N == 0 || im_h == 0 || im_w == 0) return DMM_OK;
  if (!masks || !boxes) return DMM_ERR_INVALID_ARGUMENT;
  if (!pasted && !bits && !tight) return DMM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PasteParams kp;
  kp.masks = masks; kp.boxes = boxes; kp.out = pasted; kp.bits = bits; kp.tight_ws = nullptr;
  kp.N = N; kp.M = M; kp.pad = padding; kp.im_h = im_h; kp.im_w = im_w;
  kp.words = (int)(((long long)im_h * im_w + 31) / 32);
  kp.scale = (float)((double)(M + 2 * padding) / (double)M);     // python float, rounded to fp32 by the tensor multiply
  kp.thresh = thresh;
  if (tight) {
    if (!workspace || workspace_bytes < (size_t)N * 4 * sizeof(int)) return DMM_ERR_WORKSPACE_TOO_SMALL;
    if (!aligned16(workspace)) return DMM_ERR_INVALID_ARGUMENT;
    kp.tight_ws = (int*)workspace;
    DMM_CUDA_TRY(cudaMemsetAsync(workspace, 0x7f, (size_t)N * 4 * sizeof(int), st));
  }
  if (bits) DMM_CUDA_TRY(cudaMemsetAsync(bits, 0, (size_t)N * kp.words * sizeof(uint32_t), st));
  // ~16 CTAs per SM over the whole batch, never fewer than 4 rows per CTA
  long long slabs = (16LL * kNumSMs + N - 1) / N;
  const long long max_slabs = (im_h + 3) / 4;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  kp.rows_per_cta = (int)((im_h + slabs - 1) / slabs);
  slabs = (im_h + kp.rows_per_cta - 1) / kp.rows_per_cta;
  const bool vec = im_w % 4 == 0 && (!pasted || aligned16(pasted));
  dim3 grid((unsigned)slabs, (unsigned)N);
  if (vec) paste_masks_kernel<true><<<grid, kThreads, 0, st>>>(kp);
  else paste_masks_kernel<false><<<grid, kThreads, 0, st>>>(kp);
  int rc = check_launch();
  if (rc) return rc;
  if (tight) {
    paste_tight_finalize_kernel<<<(N + 127) / 128, 128, 0, st>>>(kp.tight_ws, tight, N, im_h, im_w);
    rc = check_launch();
  }
  return rc;
}

extern "C" int dmm_paste_apply(const float* Bmat, const float* masks, const float* boxes, const int* src_index, int B, int P,
                               int O, int MS, int M, int padding, int im_h, int im_w, const int* n_prop, const int* n_tmpl,
                               const int* row_map, int O_out, int zero_fill, float* out, long long out_bstride,
                               void* stream) {
  if (B < 0 || P < 0 || O < 0 || MS < P || M <= 0 || padding < 1 || im_h < 0 || im_w < 0 || O_out < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (M + 2 * padding > kMaxMp || P > 128 || (long long)B * O_out > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;
  if (B == 0 || O_out == 0 || im_h == 0 || im_w == 0) return DMM_OK;
  if (!out || (O > 0 && P > 0 && (!Bmat || !masks || !boxes || !src_index))) return DMM_ERR_INVALID_ARGUMENT;
  PasteApplyParams kp;
  kp.coef = Bmat; kp.masks = masks; kp.boxes = boxes; kp.src = src_index; kp.n_prop = n_prop; kp.n_tmpl = n_tmpl;
  kp.row_map = row_map; kp.out = out; kp.out_bs = out_bstride;
  kp.B = B; kp.P = P; kp.O = O; kp.MS = MS; kp.O_out = O_out; kp.M = M; kp.pad = padding; kp.im_h = im_h; kp.im_w = im_w;
  kp.zero_fill = zero_fill;
  kp.scale = (float)((double)(M + 2 * padding) / (double)M);
  const long long rows_total = (long long)B * O_out;
  long long slabs = (16LL * kNumSMs + rows_total - 1) / rows_total;
  const long long max_slabs = (im_h + 3) / 4;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  kp.rows_per_cta = (int)((im_h + slabs - 1) / slabs);
  slabs = (im_h + kp.rows_per_cta - 1) / kp.rows_per_cta;
  const int Mp = M + 2 * padding;
  const size_t smem = (size_t)kMaxSrc * Mp * Mp * sizeof(float);
  const bool vec = im_w % 4 == 0 && aligned16(out) && out_bstride % 4 == 0;
  dim3 grid((unsigned)slabs, (unsigned)rows_total);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) {
    DMM_CUDA_TRY(cudaFuncSetAttribute(paste_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    paste_apply_kernel<true><<<grid, kThreads, smem, st>>>(kp);
  } else {
    DMM_CUDA_TRY(cudaFuncSetAttribute(paste_apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    paste_apply_kernel<false><<<grid, kThreads, smem, st>>>(kp);
  }
  return check_launch();
}

// ---- K10 backward: gradient w.r.t. the assignment coefficients ---------------------------------------------------------
// out[b, row(o)] = sum_p Bm[b,o,p] * paste(det src[b,p])  =>  g_Bm[b,o,p] = < g_out[b, row(o)], paste(det src[b,p]) >  over the
// detection's (clipped) box, for the entries the solver selected (sel[b,o,p] != 0: Bm = R * logic, match_model.py:128-130).
// The mask-head outputs themselves carry no gradient in the reference (proposals are loaded offline,
// dmm/modules/model_encoder.py), so this is the whole backward of the lazy pipeline: training no longer has to
// materialise the P pasted masks.  One CTA per (problem, template row); for every selected detection the padded 28x28
// mask sits in shared memory, threads walk the box with the forward's taps, a fixed-order block reduction gives the dot
// product: deterministic.
struct PasteApplyBwdParams {
  const float* gout;     // [B][O_out][im_h][im_w] (batch stride gout_bs)
  const float* sel;      // [B][O][MS]  selection mask (non-zero: gradient flows)
  const float* masks;    // [Nsrc][M][M]
  const float* boxes;    // [Nsrc][4]
  const int* src;        // [B][P]
  const int* n_prop;
  const int* n_tmpl;
  const int* row_map;    // [B][O] or NULL
  float* g_coef;         // [B][O][MS]
  long long gout_bs;
  int B, P, O, MS, O_out, M, pad, im_h, im_w;
  float scale;
};

__global__ void __launch_bounds__(kThreads) paste_apply_bwd_kernel(const PasteApplyBwdParams p) {
  __shared__ float sm[kMaxMp * kMaxMp];
  __shared__ float s_red[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.O, o = blockIdx.x % p.O;
  const int Mp = p.M + 2 * p.pad;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  float* grow = p.g_coef + ((long long)b * p.O + o) * p.MS;
  for (int c = tid; c < p.MS; c += kThreads) grow[c] = 0.f;
  const int f = o < nt ? (p.row_map ? p.row_map[(long long)b * p.O + o] : o) : -1;
  if (f < 0 || f >= p.O_out) return;                               // row not written by the forward: no gradient
  const float* g = p.gout + (long long)b * p.gout_bs + (long long)f * p.im_h * p.im_w;
  const float* srow = p.sel + ((long long)b * p.O + o) * p.MS;
  for (int c = 0; c < np; ++c) {                                   // block-uniform loop over the selected detections
    const int det = p.src[(long long)b * p.P + c];
    if (srow[c] == 0.f || det < 0) continue;
    __syncthreads();                                               // previous detection's tile / partials consumed
    const float* src = p.masks + (long long)det * p.M * p.M;
    for (int i = tid; i < Mp * Mp; i += kThreads) {
      const int y = i / Mp - p.pad, x = i % Mp - p.pad;
      sm[i] = (y >= 0 && y < p.M && x >= 0 && x < p.M) ? src[y * p.M + x] : 0.f;
    }
    const BoxI bx = expand_box(p.boxes + 4LL * det, p.scale);
    const int x0c = max(bx.x0, 0), x1c = min(bx.x1 + 1, p.im_w);
    const int y0c = max(bx.y0, 0), y1c = min(bx.y1 + 1, p.im_h);
    __syncthreads();
    float acc = 0.f;
    const int bw = x1c - x0c, bh = y1c - y0c;
    if (bw > 0 && bh > 0) {
      const float sc_y = (float)Mp / (float)bx.h, sc_x = (float)Mp / (float)bx.w;
      for (int it = tid; it < bw * bh; it += kThreads) {
        const int Y = y0c + it / bw, X = x0c + it % bw;
        const Tap ty = tap(Y - bx.y0, sc_y, Mp), tx = tap(X - bx.x0, sc_x, Mp);
        const float* r0 = sm + ty.i0 * Mp;
        const float* r1 = sm + ty.i1 * Mp;
        const float top = fmaf(tx.l0, r0[tx.i0], __fmul_rn(tx.l1, r0[tx.i1]));
        const float bot = fmaf(tx.l0, r1[tx.i0], __fmul_rn(tx.l1, r1[tx.i1]));
        const float m = fmaf(ty.l0, top, __fmul_rn(ty.l1, bot));
        acc = fmaf(g[(long long)Y * p.im_w + X], m, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t = __fadd_rn(t, s_red[w]);
      grow[c] = t;
    }
  }
}

extern "C" int dmm_paste_apply_bwd(const float* g_out, long long gout_bstride, const float* sel, const float* masks,
                                   const float* boxes, const int* src_index, int B, int P, int O, int MS, int M, int padding,
                                   int im_h, int im_w, const int* n_prop, const int* n_tmpl, const int* row_map, int O_out,
                                   float* g_Bmat, void* stream) {
  if (B < 0 || P < 0 || O < 0 || MS < P || M <= 0 || padding < 1 || im_h < 0 || im_w < 0 || O_out < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (M + 2 * padding > kMaxMp || P > 128) return DMM_ERR_UNSUPPORTED_SHAPE;
  if (B == 0 || O == 0) return DMM_OK;
  if (!g_Bmat || (P > 0 && im_h * im_w > 0 && (!g_out || !sel || !masks || !boxes || !src_index))) return DMM_ERR_INVALID_ARGUMENT;
  PasteApplyBwdParams kp;
  kp.gout = g_out; kp.gout_bs = gout_bstride; kp.sel = sel; kp.masks = masks; kp.boxes = boxes; kp.src = src_index;
  kp.n_prop = n_prop; kp.n_tmpl = n_tmpl; kp.row_map = row_map; kp.g_coef = g_Bmat;
  kp.B = B; kp.P = P; kp.O = O; kp.MS = MS; kp.O_out = O_out; kp.M = M; kp.pad = padding; kp.im_h = im_h; kp.im_w = im_w;
  kp.scale = (float)((double)(M + 2 * padding) / (double)M);
  cudaStream_t st = (cudaStream_t)stream;
  paste_apply_bwd_kernel<<<(unsigned)((long long)B * O), kThreads, 0, st>>>(kp);
  return check_launch();
}

extern "C" int dmm_box_nms(const float* boxes, const float* scores, const int* n_boxes, int F, int n_max, float thresh,
                           int max_keep, long long* keep, int* n_keep, void* stream) {
  if (F < 0 || n_max < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (n_max > kNmsMax) return DMM_ERR_UNSUPPORTED_SHAPE;
  if (F == 0) return DMM_OK;
  if (!keep || !n_keep || (n_max > 0 && (!boxes || !scores))) return DMM_ERR_INVALID_ARGUMENT;
  if (n_max > 0 && !aligned16(boxes)) return DMM_ERR_INVALID_ARGUMENT;
  box_nms_kernel<<<F, kNmsMax, 0, (cudaStream_t)stream>>>(boxes, scores, n_boxes, n_max, thresh, max_keep, keep, n_keep);
  return check_launch();
}
