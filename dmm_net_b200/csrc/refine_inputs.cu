// K6 -- decoder mask-input pyramid (forward + backward), K7 -- merged label map.
//
// Reference K6: dmm/modules/trainer.py:256-263 and dmm/modules/evaluator.py:187-194.  Per object t the reference builds
//   prev_m_inst = cat(prev_mask[:,t], ref_mask[:,t], init_pred_inst[:,t]) -> [B,3,H,W]
// and applies nn.MaxPool2d((2,2), ceil_mode=True) 1+L times, keeping the last L results (windows 4, 8, ... 2^(L+1)).
// Chained ceil-mode 2x2 pools are a max over the 2^k x 2^k window clipped to the image, so one pass over the three
// [B,O,H,W] tensors produces every level of every object: a CTA reads a 64x64 pixel tile once (16 KB, four LDG.128 per
// thread in flight) and reduces it hierarchically on chip.  HBM-bound: 3*B*O*HW*4 bytes read, ~8 % of that written.
// max is exact, so the forward is bit-equal to the reference (NaN propagates like ATen's `val > maxval || isnan(val)`).
// Backward: each stage routes the gradient to the first maximum of its 2x2 window in row-major order
// (max_pool2d_with_indices), and a level's gradient is (its own cotangent) + (what the next level routes into it): the
// kernel recomputes the (value, arg-max pixel) hierarchy of the tile and accumulates top-down in that same two-operand
// order, so the result is deterministic and bit-equal to autograd's.
//
// Reference K7: dmm/modules/evaluator.py:139-145.  label = argmax over [1 - max_o m_o, m_0, ..., m_{n-1}] per pixel
// (first maximum wins), n = number of valid templates of the video.
#include "common.cuh"

namespace dmm {
namespace {

constexpr int kTile = 64;          // input pixels per tile side; supports windows up to 64 (L <= 5)
constexpr int kThreads = 256;      // 16 x 16 threads, a 4x4 pixel patch each
constexpr int kMaxLevels = 5;

struct PyrParams {
  const float* in[3];              // prev, ref, init: [B][O][H][W]
  long long in_bs[3];              // batch strides (elements); object stride is H*W
  float* out[kMaxLevels];          // level k: [O][B][3][hk][wk]
  const float* gout[kMaxLevels];   // backward: cotangents, same layout (NULL = zero)
  float* gin[3];                   // backward: [B][O][H][W] dense (batch stride O*H*W), NULL = not wanted
  int hk[kMaxLevels], wk[kMaxLevels];
  int B, O, H, W, L;
  int tiles_x, tiles_y;
};

// ATen max_pool2d update rule: take the later value if it is larger or NaN.
__device__ __forceinline__ bool takes(float cur, float nxt) { return nxt > cur || nxt != nxt; }

struct VI { float v; int i; };
__device__ __forceinline__ VI fold(VI a, VI b) { return takes(a.v, b.v) ? b : a; }
// Forward needs values only: max.NaN.f32 is ATen's rule in one instruction (NaN wins; the larger value otherwise).  The
// forward kernel was issue-bound (ncu: 80 % issue slots busy at 64 % DRAM) on the compare/select pairs of `fold`.
__device__ __forceinline__ float vmax(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

template <bool VEC>
__device__ __forceinline__ void load_patch(const float* plane, int H, int W, int y0, int x0, float (&v)[4][4]) {
  const float ninf = __int_as_float(0xff800000);
  if (VEC && y0 + 3 < H && x0 + 3 < W) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 q = ld_stream_f4(plane + (long long)(y0 + r) * W + x0);
      v[r][0] = q.x; v[r][1] = q.y; v[r][2] = q.z; v[r][3] = q.w;
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        v[r][c] = (y0 + r < H && x0 + c < W) ? ld_stream_f1(plane + (long long)(y0 + r) * W + x0 + c) : ninf;
  }
}

// Two chained 2x2 pools over a 4x4 patch; idx = pixel index inside the 64x64 tile.
__device__ __forceinline__ VI pool_patch(const float (&v)[4][4], int ty, int tx) {
  VI q[2][2];
#pragma unroll
  for (int sy = 0; sy < 2; ++sy)
#pragma unroll
    for (int sx = 0; sx < 2; ++sx) {
      const int py = 4 * ty + 2 * sy, px = 4 * tx + 2 * sx;
      VI m = {v[2 * sy][2 * sx], py * kTile + px};
      m = fold(m, VI{v[2 * sy][2 * sx + 1], py * kTile + px + 1});
      m = fold(m, VI{v[2 * sy + 1][2 * sx], (py + 1) * kTile + px});
      m = fold(m, VI{v[2 * sy + 1][2 * sx + 1], (py + 1) * kTile + px + 1});
      q[sy][sx] = m;
    }
  return fold(fold(fold(q[0][0], q[0][1]), q[1][0]), q[1][1]);
}

// offsets of the level arrays inside the per-tile scratch: 256 + 64 + 16 + 4 + 1
__device__ __forceinline__ int lvl_off(int k) { return k == 0 ? 0 : k == 1 ? 256 : k == 2 ? 320 : k == 3 ? 336 : 340; }

template <bool VEC, bool BWD>
__global__ void __launch_bounds__(kThreads) mask_pyramid_kernel(const PyrParams p) {
  __shared__ float sv[341];
  __shared__ int si[BWD ? 341 : 1];
  __shared__ float sg[BWD ? 341 : 1];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  long long t = blockIdx.x;
  const int tile_x = (int)(t % p.tiles_x); t /= p.tiles_x;
  const int tile_y = (int)(t % p.tiles_y); t /= p.tiles_y;
  const int c = (int)(t % 3); t /= 3;
  const int b = (int)(t % p.B);
  const int o = (int)(t / p.B);
  if (BWD && p.gin[c] == nullptr) return;
  const float* plane = p.in[c] + (long long)b * p.in_bs[c] + (long long)o * p.H * p.W;
  const int y0 = tile_y * kTile + 4 * ty, x0 = tile_x * kTile + 4 * tx;
  float v[4][4];
  load_patch<VEC>(plane, p.H, p.W, y0, x0, v);
  if constexpr (BWD) {
    const VI m0 = pool_patch(v, ty, tx);
    sv[tid] = m0.v;
    si[tid] = m0.i;
  } else {
    float m = v[0][0];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) m = vmax(m, v[r][c]);
    sv[tid] = m;
  }
  __syncthreads();
  // levels 1.. : 64, 16, 4, 1 cells, each the row-major fold of a 2x2 block of the level below
  for (int k = 1; k < p.L; ++k) {
    const int n = 16 >> k;                      // cells per side at level k
    if (!BWD && tid < n * n) {
      const int cy = tid / n, cx = tid % n, lo = lvl_off(k - 1), ns = 2 * n;
      sv[lvl_off(k) + tid] = vmax(vmax(sv[lo + (2 * cy) * ns + 2 * cx], sv[lo + (2 * cy) * ns + 2 * cx + 1]),
                                  vmax(sv[lo + (2 * cy + 1) * ns + 2 * cx], sv[lo + (2 * cy + 1) * ns + 2 * cx + 1]));
    }
    if (BWD && tid < n * n) {
      const int cy = tid / n, cx = tid % n, lo = lvl_off(k - 1), ns = 2 * n;
      VI m = {sv[lo + (2 * cy) * ns + 2 * cx], BWD ? si[lo + (2 * cy) * ns + 2 * cx] : 0};
      m = fold(m, VI{sv[lo + (2 * cy) * ns + 2 * cx + 1], BWD ? si[lo + (2 * cy) * ns + 2 * cx + 1] : 0});
      m = fold(m, VI{sv[lo + (2 * cy + 1) * ns + 2 * cx], BWD ? si[lo + (2 * cy + 1) * ns + 2 * cx] : 0});
      m = fold(m, VI{sv[lo + (2 * cy + 1) * ns + 2 * cx + 1], BWD ? si[lo + (2 * cy + 1) * ns + 2 * cx + 1] : 0});
      sv[lvl_off(k) + tid] = m.v;
      if (BWD) si[lvl_off(k) + tid] = m.i;
    }
    __syncthreads();
  }
  const long long plane_id = ((long long)o * p.B + b) * 3 + c;
  if constexpr (!BWD) {
    // level 0 from registers, the rest from shared memory; a cell exists iff its first pixel is inside the image
    for (int k = 0; k < p.L; ++k) {
      const int n = 16 >> k;
      if (tid < n * n) {
        const int cy = tid / n, cx = tid % n;
        const int gy = tile_y * n + cy, gx = tile_x * n + cx;
        if (gy < p.hk[k] && gx < p.wk[k])
          p.out[k][(plane_id * p.hk[k] + gy) * p.wk[k] + gx] = sv[lvl_off(k) + tid];
      }
    }
  } else {
  // backward: top-down accumulation  g_k[cell] = gout_k[cell] + (cell holds its parent's arg-max ? g_{k+1}[parent] : 0)
  for (int k = p.L - 1; k >= 0; --k) {
    const int n = 16 >> k;
    if (tid < n * n) {
      const int cy = tid / n, cx = tid % n;
      const int gy = tile_y * n + cy, gx = tile_x * n + cx;
      float g = 0.f;
      if (p.gout[k] && gy < p.hk[k] && gx < p.wk[k]) g = p.gout[k][(plane_id * p.hk[k] + gy) * p.wk[k] + gx];
      if (k + 1 < p.L) {
        const int par = lvl_off(k + 1) + (cy >> 1) * (n >> 1) + (cx >> 1);
        if (si[par] == si[lvl_off(k) + tid]) g = __fadd_rn(g, sg[par]);
      }
      sg[lvl_off(k) + tid] = g;
    }
    __syncthreads();
  }
  const float g0 = sg[tid];
  const int arg = si[tid];
  float* gplane = p.gin[c] + ((long long)b * p.O + o) * p.H * p.W;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = y0 + r;
    if (y >= p.H) break;
    float w[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) w[cc] = ((4 * ty + r) * kTile + 4 * tx + cc == arg) ? g0 : 0.f;
    if (VEC && x0 + 3 < p.W) {
      st_stream_f4(gplane + (long long)y * p.W + x0, make_float4(w[0], w[1], w[2], w[3]));
    } else {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        if (x0 + cc < p.W) gplane[(long long)y * p.W + x0 + cc] = w[cc];
    }
  }
  }
}

// ---- K7 ------------------------------------------------------------------------------------------------------
// Inputs are sigmoid outputs (NaN-free).  torch.max(0) keeps the first maximum: the running best only moves on a strict
// `>`, and the background (index 0) wins ties against every object.
template <bool VEC>
__global__ void __launch_bounds__(256) merge_labels_kernel(const float* __restrict__ masks, long long bstride, int B, int O,
                                                           int HW, const int* __restrict__ n_valid,
                                                           unsigned char* __restrict__ label, int per_b_blocks) {
  const int b = blockIdx.x / per_b_blocks, blk = blockIdx.x % per_b_blocks;
  const int n = n_valid ? clampi(n_valid[b], 0, O) : O;
  const float* m = masks + (long long)b * bstride;
  unsigned char* out = label + (long long)b * HW;
  constexpr int step = VEC ? 4 : 1;
  const int nq = (HW + step - 1) / step;
  for (int q = blk * 256 + threadIdx.x; q < nq; q += per_b_blocks * 256) {
    float best[step];
    int arg[step];
#pragma unroll
    for (int j = 0; j < step; ++j) { best[j] = 0.f; arg[j] = 0; }
    for (int o0 = 0; o0 < n; o0 += 4) {                     // four objects' loads in flight, compared in object order
      float v[4][step];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int o = min(o0 + k, n - 1);                    // clamped (re-reads the last object; never compared)
        if constexpr (VEC) {
          const float4 f = ld_stream_f4(m + (long long)o * HW + 4 * q);
          v[k][0] = f.x; v[k][1] = f.y; v[k][2] = f.z; v[k][3] = f.w;
        } else {
          v[k][0] = ld_stream_f1(m + (long long)o * HW + q);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int o = o0 + k;
        if (o < n) {
#pragma unroll
          for (int j = 0; j < step; ++j)
            if (o == 0 || v[k][j] > best[j]) { best[j] = v[k][j]; arg[j] = o + 1; }
        }
      }
    }
    unsigned char r[step];
#pragma unroll
    for (int j = 0; j < step; ++j) {
      const float bg = __fsub_rn(1.f, best[j]);               // 1 - max_o (evaluator.py:141)
      r[j] = (unsigned char)((n > 0 && best[j] > bg) ? arg[j] : 0);
    }
    if constexpr (VEC) {
      *reinterpret_cast<uchar4*>(out + 4 * q) = make_uchar4(r[0], r[1], r[2], r[3]);
    } else {
      out[q] = r[0];
    }
  }
}

inline bool aligned16(const void* q) { return ((uintptr_t)q & 15u) == 0; }

int fill_params(PyrParams& kp, const float* prev, long long prev_bs, const float* ref, long long ref_bs, const float* init,
                long long init_bs, int B, int O, int H, int W, int L) {
  if (B < 0 || O < 0 || H < 0 || W < 0 || L < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (L > kMaxLevels) return DMM_ERR_UNSUPPORTED_SHAPE;
  kp.in[0] = prev; kp.in[1] = ref; kp.in[2] = init;
  kp.in_bs[0] = prev_bs; kp.in_bs[1] = ref_bs; kp.in_bs[2] = init_bs;
  kp.B = B; kp.O = O; kp.H = H; kp.W = W; kp.L = L;
  for (int k = 0; k < kMaxLevels; ++k) {
    const int win = 4 << k;
    kp.hk[k] = (H + win - 1) / win; kp.wk[k] = (W + win - 1) / win;
    kp.out[k] = nullptr; kp.gout[k] = nullptr;
  }
  kp.gin[0] = kp.gin[1] = kp.gin[2] = nullptr;
  kp.tiles_x = (W + kTile - 1) / kTile; kp.tiles_y = (H + kTile - 1) / kTile;
  return DMM_OK;
}

}  // namespace
}  // namespace dmm

using namespace dmm;

extern "C" int dmm_mask_pyramid_level_size(int H, int W, int level, int* h_out, int* w_out) {
  if (H < 0 || W < 0 || level < 0 || level >= kMaxLevels || !h_out || !w_out) return DMM_ERR_INVALID_ARGUMENT;
  const int win = 4 << level;
  *h_out = (H + win - 1) / win; *w_out = (W + win - 1) / win;
  return DMM_OK;
}

extern "C" int dmm_mask_pyramid(const float* prev, long long prev_bstride, const float* ref, long long ref_bstride,
                                const float* init, long long init_bstride, int B, int O, int H, int W, int L,
                                float* const* out_levels, void* stream) {
  PyrParams kp;
  int rc = fill_params(kp, prev, prev_bstride, ref, ref_bstride, init, init_bstride, B, O, H, W, L);
  if (rc) return rc;
  if (B == 0 || O == 0 || H == 0 || W == 0 || L == 0) return DMM_OK;
  if (!prev || !ref || !init || !out_levels) return DMM_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < L; ++k) {
    if (!out_levels[k]) return DMM_ERR_INVALID_ARGUMENT;
    kp.out[k] = out_levels[k];
  }
  const long long blocks = (long long)kp.tiles_x * kp.tiles_y * 3 * B * O;
  if (blocks > 0x7fffffffLL) return DMM_ERR_UNSUPPORTED_SHAPE;
  bool vec = W % 4 == 0;
  for (int c = 0; c < 3; ++c) vec = vec && aligned16(kp.in[c]) && kp.in_bs[c] % 4 == 0;
  if (vec) mask_pyramid_kernel<true, false><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(kp);
  else mask_pyramid_kernel<false, false><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(kp);
  return check_launch();
}

extern "C" int dmm_mask_pyramid_bwd(const float* const* g_out_levels, const float* prev, long long prev_bstride,
                                    const float* ref, long long ref_bstride, const float* init, long long init_bstride,
                                    int B, int O, int H, int W, int L, float* g_prev, float* g_ref, float* g_init,
                                    void* stream) {
  PyrParams kp;
  int rc = fill_params(kp, prev, prev_bstride, ref, ref_bstride, init, init_bstride, B, O, H, W, L);
  if (rc) return rc;
  if (B == 0 || O == 0 || H == 0 || W == 0) return DMM_OK;
  if (!g_prev && !g_ref && !g_init) return DMM_OK;
  if (!g_out_levels && L > 0) return DMM_ERR_INVALID_ARGUMENT;
  if ((g_prev && !prev) || (g_ref && !ref) || (g_init && !init)) return DMM_ERR_INVALID_ARGUMENT;
  for (int k = 0; k < L; ++k) kp.gout[k] = g_out_levels[k];
  kp.gin[0] = g_prev; kp.gin[1] = g_ref; kp.gin[2] = g_init;
  const long long blocks = (long long)kp.tiles_x * kp.tiles_y * 3 * B * O;
  if (blocks > 0x7fffffffLL) return DMM_ERR_UNSUPPORTED_SHAPE;
  bool vec = W % 4 == 0;
  for (int c = 0; c < 3; ++c)
    if (kp.gin[c]) vec = vec && aligned16(kp.in[c]) && kp.in_bs[c] % 4 == 0 && aligned16(kp.gin[c]);
  if (vec) mask_pyramid_kernel<true, true><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(kp);
  else mask_pyramid_kernel<false, true><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(kp);
  return check_launch();
}

extern "C" int dmm_merge_labels(const float* masks, long long bstride, int B, int O, int HW, const int* n_valid,
                                unsigned char* label, void* stream) {
  if (B < 0 || O < 0 || HW < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || HW == 0) return DMM_OK;
  if (!label || (O > 0 && !masks)) return DMM_ERR_INVALID_ARGUMENT;
  if (O > 254) return DMM_ERR_UNSUPPORTED_SHAPE;
  const bool vec = HW % 4 == 0 && aligned16(masks) && bstride % 4 == 0 && ((uintptr_t)label & 3u) == 0;
  const int nq = vec ? HW / 4 : HW;
  long long per_b = (nq + 1023) / 1024;                     // four items per thread: every thread does the same work
  if (per_b < 1) per_b = 1;
  const long long blocks = per_b * B;
  if (blocks > 0x7fffffffLL) return DMM_ERR_UNSUPPORTED_SHAPE;
  if (vec) merge_labels_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(masks, bstride, B, O, HW, n_valid, label, (int)per_b);
  else merge_labels_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(masks, bstride, B, O, HW, n_valid, label, (int)per_b);
  return check_launch();
}
