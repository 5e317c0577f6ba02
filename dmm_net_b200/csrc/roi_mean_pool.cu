// K5 -- proposal-feature pooling: legacy ROIAlign(14x14, sampling_ratio 2) on 4 feature levels + spatial mean.
//
// Reference: dmm/modules/feature_extractor.py:11-52 over maskrcnn_benchmark's Pooler/ROIAlign (un-vendored fork,
// no pinned commit: parity is against torchvision.ops.roi_align(aligned=False), the same legacy arithmetic).
//
// The mean over the 14x14 bins of 2x2 bilinear samples is linear in the feature map and SEPARABLE:
//   out[r, l*C + c] = sum_y sum_x wy[r,l,y] * wx[r,l,x] * F_l[b_r, c, y, x]
// (the "sample outside [-1, size] contributes 0" rule factorises per axis).  So instead of 196*4 gathers of 4 taps
// per channel the kernel builds the two 1-D weight vectors once per (ROI, level) and contracts the ROI's window
// with coalesced row reads; fp32 FFMA keeps the 1e-4 bar on the downstream cosine.
#include "common.cuh"

namespace dmm {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRes = 14;       // feature_extractor.py:15
constexpr int kSamp = 2;       // feature_extractor.py:14
constexpr int kNS = kRes * kSamp;
constexpr int kTabCap = 2048;  // window cells per forward table chunk

struct PoolParams {
  const float* feat[4];
  float* gfeat[4];
  int H[4], W[4];
  int N, C, R;
  int level_skip;      // bit l set: level l is computed elsewhere (tensor-core path), this kernel leaves it alone
  const float* rois;   // [R][5]
  float* out;          // [R][4*C]
  const float* gout;
};

// weight of index i on an axis of length `size` for the ROI interval [lo, hi] (image px) at `scale`
__device__ __forceinline__ float axis_weight(int i, float lo, float hi, int size, float scale) {
  const float start = lo * scale;
  const float len = fmaxf(hi * scale - start, 1.f);
  const float bin = len / (float)kRes;
  float w = 0.f;
#pragma unroll 4
  for (int s = 0; s < kNS; ++s) {
    const int bidx = s / kSamp, sidx = s - bidx * kSamp;
    float t = start + (float)bidx * bin + ((float)sidx + 0.5f) * bin / (float)kSamp;
    if (t < -1.f || t > (float)size) continue;
    if (t <= 0.f) t = 0.f;
    int l = (int)t, h;
    if (l >= size - 1) { l = h = size - 1; t = (float)l; } else h = l + 1;
    const float fr = t - (float)l;
    if (i == l) w += 1.f - fr;
    if (i == h) w += fr;
  }
  return w * (1.f / (float)kNS);
}

__device__ __forceinline__ void axis_range(float lo, float hi, int size, float scale, int& a, int& b) {
  const float start = lo * scale;
  const float len = fmaxf(hi * scale - start, 1.f);
  const float t0 = start + 0.5f * (len / (float)kRes) / (float)kSamp;
  const float t1 = start + len;
  a = clampi((int)floorf(fmaxf(t0, 0.f)) - 1, 0, size - 1);
  b = clampi((int)floorf(t1) + 1, 0, size - 1);
}

// TABLE = true: the ROI window is flattened into a shared-memory table of (element offset, wy*wx) pairs built once
// per (ROI, level) and reused by all channels: lanes walk the window linearly (full lane utilisation even for the
// 3..8 pixel wide windows of the coarse levels, where a lane-per-column loop idles 75-90 % of the warp).
// TABLE = false: same arithmetic with on-the-fly index math, for feature maps too large for the table.
template <bool BWD, bool TABLE>
__global__ void __launch_bounds__(kThreads) roi_mean_pool_kernel(const PoolParams p) {
  extern __shared__ float wsm[];  // wy[H] wx[W] | table: off[hh*ww] (int), w[hh*ww] (float)
  const int r = blockIdx.x, l = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((p.level_skip >> l) & 1) return;
  const int H = p.H[l], W = p.W[l];
  const float scale = 0.25f / (float)(1 << l);  // 1/4, 1/8, 1/16, 1/32 (feature_extractor.py:13)
  const float* roi = p.rois + (long long)r * 5;
  const int n = (int)roi[0];
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  float* wy = wsm;
  float* wx = wsm + H;
  for (int i = tid; i < H; i += kThreads) wy[i] = axis_weight(i, y1, y2, H, scale);
  for (int i = tid; i < W; i += kThreads) wx[i] = axis_weight(i, x1, x2, W, scale);
  int ya, yb, xa, xb;
  axis_range(y1, y2, H, scale, ya, yb);
  axis_range(x1, x2, W, scale, xa, xb);
  const int ww = xb - xa + 1, hh = yb - ya + 1, cnt = ww * hh;
  // Forward: the table holds at most kTabCap window cells at a time (16 KB: 8 CTAs per SM instead of the 3 that a table
  // sized for the whole level-0 map allowed); larger windows are walked in chunks.  Backward keeps the full-map table.
  const int tab_cap = BWD ? H * W : kTabCap;
  int* toff = reinterpret_cast<int*>(wsm + H + W);
  float* tw = wsm + H + W + (TABLE ? tab_cap : 0);
  __syncthreads();
  const bool valid = n >= 0 && n < p.N;
  if (TABLE && BWD) {
    for (int i = tid; i < cnt; i += kThreads) {
      const int y = ya + i / ww, x = xa + i % ww;
      toff[i] = y * W + x;
      tw[i] = wy[y] * wx[x];
    }
    __syncthreads();
  }
  if (!BWD) {
    float* o = p.out + (long long)r * 4 * p.C + (long long)l * p.C;
    if (TABLE) {
      for (int base = 0; base == 0 || base < cnt; base += kTabCap) {
        const int m = min(kTabCap, cnt - base);
        if (base > 0) __syncthreads();                       // previous chunk consumed by every warp
        for (int i = tid; i < m; i += kThreads) {
          const int idx = base + i, y = ya + idx / ww, x = xa + idx % ww;
          toff[i] = y * W + x;
          tw[i] = wy[y] * wx[x];
        }
        __syncthreads();
        for (int c = warp; c < p.C; c += kWarps) {
          float acc = 0.f;
          if (valid) {
            const float* f = p.feat[l] + ((long long)n * p.C + c) * H * W;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;    // four gathers in flight per lane: the loop is L2-latency bound
            int i = lane;
            for (; i + 96 < m; i += 128) {
              const float v0 = __ldg(f + toff[i]), v1 = __ldg(f + toff[i + 32]);
              const float v2 = __ldg(f + toff[i + 64]), v3 = __ldg(f + toff[i + 96]);
              a0 = fmaf(tw[i], v0, a0); a1 = fmaf(tw[i + 32], v1, a1);
              a2 = fmaf(tw[i + 64], v2, a2); a3 = fmaf(tw[i + 96], v3, a3);
            }
            for (; i < m; i += 32) a0 = fmaf(tw[i], __ldg(f + toff[i]), a0);
            acc = (a0 + a1) + (a2 + a3);
          }
          acc = warp_sum(acc);
          if (lane == 0) o[c] = base == 0 ? acc : o[c] + acc;
        }
      }
    } else {
      for (int c = warp; c < p.C; c += kWarps) {
        float acc = 0.f;
        if (valid) {
          const float* f = p.feat[l] + ((long long)n * p.C + c) * H * W;
          for (int y = ya; y <= yb; ++y) {
            const float wyv = wy[y];
            if (wyv == 0.f) continue;
            float rowacc = 0.f;
            for (int x = xa + lane; x <= xb; x += 32) rowacc = fmaf(wx[x], f[y * W + x], rowacc);
            acc = fmaf(wyv, rowacc, acc);
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) o[c] = acc;
      }
    }
  } else {
    if (!valid) return;
    const float* go = p.gout + (long long)r * 4 * p.C + (long long)l * p.C;
    for (int c = warp; c < p.C; c += kWarps) {
      const float g = go[c];
      if (g == 0.f) continue;
      float* f = p.gfeat[l] + ((long long)n * p.C + c) * H * W;
      for (int i = lane; i < cnt; i += 32) {
        int off;
        float wgt;
        if (TABLE) { off = toff[i]; wgt = tw[i]; }
        else { const int y = ya + i / ww, x = xa + i % ww; off = y * W + x; wgt = wy[y] * wx[x]; }
        if (wgt != 0.f) atomicAdd(f + off, g * wgt);
      }
    }
  }
}

}  // namespace
}  // namespace dmm

using namespace dmm;

static int fill(PoolParams& kp, const int Hl[4], const int Wl[4], int N, int C, const float* rois, int R, size_t& smem,
                size_t& table_smem) {
  if (N < 0 || C < 0 || R < 0 || !Hl || !Wl) return DMM_ERR_INVALID_ARGUMENT;
  smem = 0;
  for (int l = 0; l < 4; ++l) {
    if (Hl[l] <= 0 || Wl[l] <= 0) return DMM_ERR_INVALID_ARGUMENT;
    kp.H[l] = Hl[l]; kp.W[l] = Wl[l];
    const size_t need = (size_t)(Hl[l] + Wl[l]) * sizeof(float);
    if (need > smem) smem = need;
    kp.feat[l] = nullptr; kp.gfeat[l] = nullptr;
  }
  if (smem > 48 * 1024) return DMM_ERR_UNSUPPORTED_SHAPE;
  size_t tab = 0;
  for (int l = 0; l < 4; ++l) {
    const size_t need = (size_t)(Hl[l] + Wl[l]) * sizeof(float) + (size_t)Hl[l] * Wl[l] * 8;
    if (need > tab) tab = need;
  }
  table_smem = tab <= 200 * 1024 ? tab : 0;   // backward; 0: feature maps too large for the window table -> index math path
  kp.N = N; kp.C = C; kp.R = R; kp.rois = rois; kp.out = nullptr; kp.gout = nullptr; kp.level_skip = 0;
  return DMM_OK;
}

namespace dmm {
size_t roi_pool_tc_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R);
int roi_pool_tc_try_launch(const float* const feat[4], const int Hl[4], const int Wl[4], int N, int C, const float* rois, int R,
                           float* out, void* workspace, size_t workspace_bytes, int* level_mask, cudaStream_t st);
}  // namespace dmm

extern "C" size_t dmm_roi_mean_pool_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R) {
  if (!Hl || !Wl) return 0;
  return roi_pool_tc_workspace_bytes(Hl, Wl, N, C, R);
}

// impl: 0 = auto (tensor-core path for every level inside its envelope when a workspace is given, SIMT kernel for the
// rest), 1 = SIMT kernel only, 2 = tensor-core path required for at least one level (DMM_ERR_UNSUPPORTED_SHAPE otherwise).
extern "C" int dmm_roi_mean_pool(const float* const feat[4], const int Hl[4], const int Wl[4], int N, int C,
                                 const float* rois, int R, float* out, void* workspace, size_t workspace_bytes, int impl,
                                 void* stream) {
  PoolParams kp; size_t smem, tsmem;
  int rc = fill(kp, Hl, Wl, N, C, rois, R, smem, tsmem);
  if (rc) return rc;
  if (impl < 0 || impl > 2) return DMM_ERR_INVALID_ARGUMENT;
  if (R == 0 || C == 0) return DMM_OK;
  if (!feat || !rois || !out) return DMM_ERR_INVALID_ARGUMENT;
  for (int l = 0; l < 4; ++l) { if (!feat[l]) return DMM_ERR_INVALID_ARGUMENT; kp.feat[l] = feat[l]; }
  kp.out = out;
  (void)tsmem;
  int tc_mask = 0;
  if (impl != 1) {
    rc = roi_pool_tc_try_launch(feat, Hl, Wl, N, C, rois, R, out, workspace, workspace_bytes, &tc_mask, (cudaStream_t)stream);
    if (rc > 0) return rc;
    if (rc < 0) tc_mask = 0;
    if (impl == 2 && tc_mask == 0) return DMM_ERR_UNSUPPORTED_SHAPE;
  }
  if (tc_mask == 15) return DMM_OK;
  kp.level_skip = tc_mask;
  const size_t fsmem = smem + (size_t)kTabCap * 8;            // wy, wx + one table chunk: ~17 KB, 8 CTAs per SM
  if (fsmem > 48 * 1024)
    DMM_CUDA_TRY(cudaFuncSetAttribute(roi_mean_pool_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
  roi_mean_pool_kernel<false, true><<<dim3(R, 4), kThreads, fsmem, (cudaStream_t)stream>>>(kp);
  return check_launch();
}

namespace dmm {
size_t roi_pool_bwd_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R);
int roi_pool_bwd_try_launch(const float* g_out, const int Hl[4], const int Wl[4], int N, int C, const float* rois, int R,
                            float* const g_feat[4], void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace dmm

extern "C" size_t dmm_roi_mean_pool_bwd_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R) {
  if (!Hl || !Wl) return 0;
  return roi_pool_bwd_workspace_bytes(Hl, Wl, N, C, R);
}

// impl: 0 = auto (deterministic gather when a workspace is given and the shapes fit: g_feat is then fully OVERWRITTEN;
// otherwise the atomic scatter, which ACCUMULATES into g_feat -- the caller zero-initialises it), 1 = atomic scatter only,
// 2 = deterministic gather required.  *wrote_all (optional) tells which of the two happened (1: overwritten).
extern "C" int dmm_roi_mean_pool_bwd(const float* g_out, const int Hl[4], const int Wl[4], int N, int C, const float* rois, int R,
                                     float* const g_feat[4], void* workspace, size_t workspace_bytes, int impl, int* wrote_all,
                                     void* stream) {
  PoolParams kp; size_t smem, tsmem;
  int rc = fill(kp, Hl, Wl, N, C, rois, R, smem, tsmem);
  if (rc) return rc;
  if (impl < 0 || impl > 2) return DMM_ERR_INVALID_ARGUMENT;
  if (wrote_all) *wrote_all = 0;
  if (R == 0 || C == 0) return DMM_OK;
  if (!g_feat || !rois || !g_out) return DMM_ERR_INVALID_ARGUMENT;
  for (int l = 0; l < 4; ++l) { if (!g_feat[l]) return DMM_ERR_INVALID_ARGUMENT; kp.gfeat[l] = g_feat[l]; }
  kp.gout = g_out;
  if (impl != 1) {
    rc = roi_pool_bwd_try_launch(g_out, Hl, Wl, N, C, rois, R, g_feat, workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc == DMM_OK) { if (wrote_all) *wrote_all = 1; return DMM_OK; }
    if (rc > 0) return rc;
    if (impl == 2) return DMM_ERR_UNSUPPORTED_SHAPE;
  }
  if (tsmem) {
    DMM_CUDA_TRY(cudaFuncSetAttribute(roi_mean_pool_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    roi_mean_pool_kernel<true, true><<<dim3(R, 4), kThreads, tsmem, (cudaStream_t)stream>>>(kp);
  } else {
    roi_mean_pool_kernel<true, false><<<dim3(R, 4), kThreads, smem, (cudaStream_t)stream>>>(kp);
  }
  return check_launch();
}
