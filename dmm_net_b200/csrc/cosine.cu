// K2 -- pairwise cosine similarity between template and proposal features, forward and backward.
//
// Reference: dmm/utils/match_helper.py:51-64 (F.cosine_similarity over D on expanded [O,D,P] views, eps 1e-8)
// averaged over the template-feature sets (dmm/modules/match_model.py:72-76; one set in practice,
// dmm/modules/dmm_model.py:44).  ATen normalises each vector by max(||v||, eps) first, multiplies, then sums.
//
// This file holds the plain fp32 FFMA forward (general shapes, several template sets, ~3e-7 of fp64) and the backward;
// the default forward for the shapes the layer runs is the tcgen05 3xTF32 kernel in cosine_tc.cu (~3e-6 of fp64).
// One CTA per problem; the feature tiles sit in shared memory, lanes are proposals.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace dmm {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowTile = 16;
constexpr int kSmemFloats = 12288;  // 48 KB of normalised template rows

struct CosParams {
  const float* q;  // [B][T][O][D]
  const float* k;  // [B][P][D]
  int B, T, P, O, D;
  const int* n_prop;
  const int* n_tmpl;
  float eps;
  float* cos;       // [B][O][P]
  const float* g;   // bwd: [B][O][P]
  const float* cos_fwd;  // bwd, optional: the forward's output (valid as cos_t only when T == 1)
  float* gq;        // bwd: [B][T][O][D]
  float* gk;        // bwd: [B][P][D]
};

__device__ __forceinline__ float row_norm(const float* v, int D, int lane) {
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(v[d], v[d], s);
  return __fsqrt_rn(warp_sum(s));
}

// Forward, v2 (round 1: v1 made lanes walk D and re-derived every smem address per (row, d): 19.5 k warp-instructions
// per warp, 55 us at B=64).  Now LANES ARE PROPOSALS: feature tiles of kDT columns are staged in shared memory with a
// +1 padded pitch (conflict-free column walks), a warp task is (one template row or the norm row) x (32 proposals), the
// template value is an smem broadcast: 2 LDS + 1 FFMA per (row, proposal, d).  The raw dot product and the squared
// norms are accumulated in one pass; cos = dot / (max(|q|,eps) * max(|k|,eps)).
constexpr int kDT = 64;                 // feature columns per tile
constexpr int kPitch = kDT + 1;
constexpr int kMaxP = 128;              // proposals per problem (solver limit)
constexpr int kMaxTasks = (kRowTile + 1) * (kMaxP / 32);   // 68
constexpr int kTPW = (kMaxTasks + kWarps - 1) / kWarps;    // 9 accumulators per lane at most

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;   // src-size 0 zero-fills the destination
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr size_t kCosSmem = (size_t)2 * (kMaxP + kRowTile) * kPitch * sizeof(float);   // double-buffered tiles: 74.9 KB

__global__ void __launch_bounds__(kThreads) cosine_fwd_kernel(const CosParams p) {
  extern __shared__ float tile[];                 // [2][ (kMaxP + kRowTile) * kPitch ]: proposals then templates
  __shared__ float knorm2[kMaxP];
  __shared__ float qnorm[kRowTile];
  constexpr int kBuf = (kMaxP + kRowTile) * kPitch;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  const int D = p.D;
  float* cosb = p.cos + (long long)b * p.O * p.P;
  const float* kb = p.k + (long long)b * p.P * D;
  const int npb = (np + 31) / 32;                 // proposal blocks of 32 lanes
  const float invT = 1.f / (float)p.T;
  const int ntiles = (D + kDT - 1) / kDT;
  for (int i = tid; i < p.O * p.P; i += kThreads) cosb[i] = 0.f;   // padding rows / columns are defined as 0

  for (int o0 = 0; o0 < nt; o0 += kRowTile) {
    const int ro = min(kRowTile, nt - o0);
    const int ntask = (ro + 1) * npb;             // row index ro is the |k|^2 task
    float tot[kTPW];                              // sum over template sets of cos_t
#pragma unroll
    for (int i = 0; i < kTPW; ++i) tot[i] = 0.f;
    for (int t = 0; t < p.T; ++t) {
      const float* qb = p.q + ((long long)b * p.T + t) * p.O * D + (long long)o0 * D;
      // cp.async stage of feature columns [d0, d0+kDT) into buffer `buf` (zero-filled past D)
      auto stage_tile = [&](int ti, int buf) {
        float* ks = tile + buf * kBuf;
        float* qs = ks + kMaxP * kPitch;
        const int d0 = ti * kDT, dt = min(kDT, D - d0);
        for (int i = tid; i < np * kDT; i += kThreads) {
          const int r = i / kDT, c = i - r * kDT;
          cp_async_f32(ks + r * kPitch + c, kb + (long long)r * D + d0 + (c < dt ? c : 0), c < dt);
        }
        for (int i = tid; i < ro * kDT; i += kThreads) {
          const int r = i / kDT, c = i - r * kDT;
          cp_async_f32(qs + r * kPitch + c, qb + (long long)r * D + d0 + (c < dt ? c : 0), c < dt);
        }
        cp_async_commit();
      };
      float acc[kTPW];
#pragma unroll
      for (int i = 0; i < kTPW; ++i) acc[i] = 0.f;
      float qn2 = 0.f;                            // warp w accumulates |q_w|^2 when the tile has <= 8 rows
      __syncthreads();                            // buffers free (previous template set / row tile done)
      if (ntiles > 0) stage_tile(0, 0);
      for (int ti = 0; ti < ntiles; ++ti) {
        const int buf = ti & 1;
        if (ti + 1 < ntiles) { stage_tile(ti + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();                          // tile ti landed for every thread
        const float* ks = tile + buf * kBuf;
        const float* qs = ks + kMaxP * kPitch;
#pragma unroll
        for (int i = 0; i < kTPW; ++i) {
          const int task = warp + kWarps * i;
          if (task < ntask) {                     // warp-uniform
            const int r = task / npb, pb = task - r * npb;
            const int col = pb * 32 + lane;
            const float* kp = ks + (col < np ? col : 0) * kPitch;
            float a0 = 0.f, a1 = 0.f;
            if (r < ro) {
              const float* qp = qs + r * kPitch;
#pragma unroll 8
              for (int c = 0; c < kDT; c += 2) {
                a0 = fmaf(qp[c], kp[c], a0);
                a1 = fmaf(qp[c + 1], kp[c + 1], a1);
              }
            } else {
#pragma unroll 8
              for (int c = 0; c < kDT; c += 2) {
                a0 = fmaf(kp[c], kp[c], a0);
                a1 = fmaf(kp[c + 1], kp[c + 1], a1);
              }
            }
            acc[i] += a0 + a1;
          }
        }
        if (ro <= kWarps && warp < ro) {          // |q_r|^2 of row r = warp rides along (more than 8 rows: pass below)
          const float* qp = qs + warp * kPitch;
          for (int c = lane; c < kDT; c += 32) qn2 = fmaf(qp[c], qp[c], qn2);
        }
        __syncthreads();                          // tile ti consumed: its buffer may be refilled by stage_tile(ti + 2)
      }
      // ---- norms: |q_r| per row (warps own rows r = warp, warp + 8), |k_p|^2 from the norm tasks ----------------
      if (ro <= kWarps) {
        const float s = warp_sum(qn2);
        if (lane == 0 && warp < ro) qnorm[warp] = fmaxf(__fsqrt_rn(s), p.eps);
      } else {
        for (int r = warp; r < ro; r += kWarps) {
          const float v = row_norm(qb + (long long)r * D, D, lane);
          if (lane == 0) qnorm[r] = fmaxf(v, p.eps);
        }
      }
#pragma unroll
      for (int i = 0; i < kTPW; ++i) {
        const int task = warp + kWarps * i;
        if (task < ntask && task / npb == ro) {
          const int col = (task - ro * npb) * 32 + lane;
          if (col < np) knorm2[col] = acc[i];
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kTPW; ++i) {
        const int task = warp + kWarps * i;
        if (task < ntask) {
          const int r = task / npb, col = (task - r * npb) * 32 + lane;
          if (r < ro && col < np) {
            const float nk = fmaxf(__fsqrt_rn(knorm2[col]), p.eps);
            tot[i] += __fdiv_rn(acc[i], __fmul_rn(qnorm[r], nk));
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kTPW; ++i) {
      const int task = warp + kWarps * i;
      if (task < ntask) {
        const int r = task / npb, col = (task - r * npb) * 32 + lane;
        if (r < ro && col < np) cosb[(o0 + r) * p.P + col] = p.T > 1 ? tot[i] * invT : tot[i];   // feature_sim /= T
      }
    }
  }
}

// Backward.  With qh = q/nq, kh = k/nk (nq, nk clamped at eps) and w = g/T:
//   g_q[o] = ( sum_p w[o,p] kh_p  -  (sum_p w[o,p] cos_t[o,p]) * q_o/|q_o| ) / nq_o
//   g_k[p] = ( sum_{t,o} w[o,p] qh_{t,o}  -  (sum_{t,o} w[o,p] cos_t[o,p]) * k_p/|k_p| ) / nk_p
// (the clamp sits under NoGradGuard in ATen, so the norm's own gradient uses the unclamped norm; 0 for a zero vector).
__global__ void __launch_bounds__(kThreads) cosine_bwd_kernel(const CosParams p) {
  extern __shared__ float smem[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  const int D = p.D, P = p.P;
  // smem: w[kRowTile][P], ct[kRowTile][P] (cos_t of the tile), nk[P], rk[P] (raw |k|), sk[P], nq[kRowTile], rq[kRowTile], sq[kRowTile]
  float* w = smem;
  float* ct = w + kRowTile * P;
  float* nk = ct + kRowTile * P;
  float* rk = nk + P;
  float* sk = rk + P;
  float* nq = sk + P;
  float* rq = nq + kRowTile;
  float* sq = rq + kRowTile;
  const float* kb = p.k + (long long)b * P * D;
  float* gkb = p.gk + (long long)b * P * D;
  const float* gb = p.g + (long long)b * p.O * P;
  const float invT = 1.f / (float)p.T;

  for (long long i = tid; i < (long long)P * D; i += kThreads) gkb[i] = 0.f;
  for (int c = warp; c < P; c += kWarps) {
    const float r = c < np ? row_norm(kb + (long long)c * D, D, lane) : 0.f;
    if (lane == 0) { rk[c] = r; nk[c] = fmaxf(r, p.eps); }
  }
  for (int t = 0; t < p.T; ++t) {
    const float* qb = p.q + ((long long)b * p.T + t) * p.O * D;
    float* gqb = p.gq + ((long long)b * p.T + t) * p.O * D;
    for (long long i = tid; i < (long long)p.O * D; i += kThreads) gqb[i] = 0.f;
    for (int o0 = 0; o0 < nt; o0 += kRowTile) {
      const int ro = min(kRowTile, nt - o0);
      __syncthreads();
      for (int r = warp; r < ro; r += kWarps) {
        const float v = row_norm(qb + (long long)(o0 + r) * D, D, lane);
        if (lane == 0) { rq[r] = v; nq[r] = fmaxf(v, p.eps); }
      }
      for (int i = tid; i < ro * P; i += kThreads) {
        const int r = i / P, c = i - r * P;
        w[r * P + c] = c < np ? gb[(o0 + r) * P + c] * invT : 0.f;
      }
      __syncthreads();
      // phase 1: cos_t of the tile -- the forward's output when there is one template set (the only case the reference
      // ever runs, dmm_model.py:44), else recomputed (one warp per (row, proposal) pair)
      if (p.cos_fwd != nullptr && p.T == 1) {
        const float* cf = p.cos_fwd + (long long)b * p.O * P;
        for (int i = tid; i < ro * np; i += kThreads) {
          const int r = i / np, c = i - r * np;
          ct[r * P + c] = cf[(o0 + r) * P + c];
        }
      } else {
        for (int i = warp; i < ro * np; i += kWarps) {
          const int r = i / np, c = i - r * np;
          const float* qv = qb + (long long)(o0 + r) * D;
          const float* kv = kb + (long long)c * D;
          float s = 0.f;
          for (int d = lane; d < D; d += 32) s = fmaf(qv[d] / nq[r], kv[d] / nk[c], s);
          s = warp_sum(s);
          if (lane == 0) ct[r * P + c] = s;
        }
      }
      __syncthreads();
      for (int r = warp; r < ro; r += kWarps) {  // sq[r] = sum_p w*cos
        float s = 0.f;
        for (int c = lane; c < np; c += 32) s = fmaf(w[r * P + c], ct[r * P + c], s);
        s = warp_sum(s);
        if (lane == 0) sq[r] = s;
      }
      for (int c = warp; c < np; c += kWarps) {  // sk[c] = sum_o w*cos (this tile)
        float s = 0.f;
        for (int r = lane; r < ro; r += 32) s = fmaf(w[r * P + c], ct[r * P + c], s);
        s = warp_sum(s);
        if (lane == 0) sk[c] = s;
      }
      __syncthreads();
      // phase 2: each thread owns feature index d
      for (int d = tid; d < D; d += kThreads) {
        float qh[kRowTile], A[kRowTile];
#pragma unroll
        for (int r = 0; r < kRowTile; ++r) {
          qh[r] = r < ro ? qb[(long long)(o0 + r) * D + d] / nq[r] : 0.f;
          A[r] = 0.f;
        }
        const bool first = (t == 0 && o0 == 0);   // first tile stores, later tiles accumulate (no RMW round trips)
#pragma unroll 4
        for (int c = 0; c < np; ++c) {
          const float kraw = __ldg(kb + (long long)c * D + d);
          const float kh = kraw / nk[c];
          float bq = 0.f;
#pragma unroll
          for (int r = 0; r < kRowTile; ++r)
            if (r < ro) {
              const float wv = w[r * P + c];
              A[r] = fmaf(wv, kh, A[r]);
              bq = fmaf(wv, qh[r], bq);
            }
          const float unit = rk[c] > 0.f ? kraw / rk[c] : 0.f;
          const float gval = (bq - sk[c] * unit) / nk[c];
          if (first) gkb[(long long)c * D + d] = gval;
          else gkb[(long long)c * D + d] += gval;
        }
#pragma unroll
        for (int r = 0; r < kRowTile; ++r)
          if (r < ro) {
            const float qraw = qb[(long long)(o0 + r) * D + d];
            const float unit = rq[r] > 0.f ? qraw / rq[r] : 0.f;
            gqb[(long long)(o0 + r) * D + d] = (A[r] - sq[r] * unit) / nq[r];
          }
      }
    }
  }
}

}  // namespace
}  // namespace dmm

using namespace dmm;

namespace dmm {
int cosine_tc_try_launch(const float* tmpl_feat, const float* prop_feat, int B, int T, int P, int O, int D,
                         const int* n_prop, const int* n_tmpl, float eps, float* cos, cudaStream_t st);   // cosine_tc.cu
}

// Default implementation of the forward: the tcgen05 kernel wherever its envelope allows (T == 1, P <= 64, O <= 16,
// D >= 32 and a multiple of 4, 16-byte aligned features), else the fp32 FFMA kernel.  DMM_K2_IMPL=simt|tc overrides the
// default for dmm_cosine_pairwise (read once); dmm_cosine_pairwise_impl selects per call.
static int k2_default_impl() {
  static const int impl = [] {
    const char* e = getenv("DMM_K2_IMPL");
    if (e && !strcmp(e, "simt")) return DMM_COSINE_SIMT;
    if (e && !strcmp(e, "tc")) return DMM_COSINE_TC;
    return DMM_COSINE_AUTO;
  }();
  return impl;
}

static int fill_params(CosParams& kp, const float* q, const float* k, int B, int T, int P, int O, int D,
                       const int* n_prop, const int* n_tmpl, float eps) {
  if (B < 0 || T < 1 || P < 0 || O < 0 || D < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (D > kSmemFloats) return DMM_ERR_UNSUPPORTED_SHAPE;
  kp.q = q; kp.k = k; kp.B = B; kp.T = T; kp.P = P; kp.O = O; kp.D = D;
  kp.n_prop = n_prop; kp.n_tmpl = n_tmpl; kp.eps = eps;
  kp.cos = nullptr; kp.g = nullptr; kp.cos_fwd = nullptr; kp.gq = nullptr; kp.gk = nullptr;
  return DMM_OK;
}

extern "C" int dmm_cosine_pairwise_impl(const float* tmpl_feat, const float* prop_feat, int B, int T, int P, int O, int D,
                                        const int* n_prop, const int* n_tmpl, float eps, float* cos, int impl,
                                        void* stream) {
  CosParams kp;
  int rc = fill_params(kp, tmpl_feat, prop_feat, B, T, P, O, D, n_prop, n_tmpl, eps);
  if (rc) return rc;
  if (impl != DMM_COSINE_AUTO && impl != DMM_COSINE_SIMT && impl != DMM_COSINE_TC) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || P == 0 || O == 0) return DMM_OK;
  if (!tmpl_feat || !prop_feat || !cos) return DMM_ERR_INVALID_ARGUMENT;
  kp.cos = cos;
  if (P > kMaxP) return DMM_ERR_UNSUPPORTED_SHAPE;
  if (impl != DMM_COSINE_SIMT) {
    rc = D >= 32 ? cosine_tc_try_launch(tmpl_feat, prop_feat, B, T, P, O, D, n_prop, n_tmpl, eps, cos, (cudaStream_t)stream) : -1;
    if (rc >= 0) return rc;            // launched (or a CUDA error); -1: outside the tensor-core kernel's envelope
    if (impl == DMM_COSINE_TC) return DMM_ERR_UNSUPPORTED_SHAPE;
  }
  DMM_CUDA_TRY(cudaFuncSetAttribute(cosine_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCosSmem));
  cosine_fwd_kernel<<<B, kThreads, kCosSmem, (cudaStream_t)stream>>>(kp);
  return check_launch();
}

extern "C" int dmm_cosine_pairwise(const float* tmpl_feat, const float* prop_feat, int B, int T, int P, int O, int D,
                                   const int* n_prop, const int* n_tmpl, float eps, float* cos, void* stream) {
  int impl = k2_default_impl();
  if (impl == DMM_COSINE_TC) impl = DMM_COSINE_AUTO;   // the env override never turns an unsupported shape into an error
  return dmm_cosine_pairwise_impl(tmpl_feat, prop_feat, B, T, P, O, D, n_prop, n_tmpl, eps, cos, impl, stream);
}

extern "C" int dmm_cosine_pairwise_bwd(const float* g_cos, const float* cos_fwd, const float* tmpl_feat,
                                       const float* prop_feat, int B, int T, int P, int O, int D, const int* n_prop,
                                       const int* n_tmpl, float eps, float* g_tmpl_feat, float* g_prop_feat,
                                       void* stream) {
  CosParams kp;
  int rc = fill_params(kp, tmpl_feat, prop_feat, B, T, P, O, D, n_prop, n_tmpl, eps);
  if (rc) return rc;
  if (B == 0 || P == 0 || O == 0 || D == 0) return DMM_OK;
  if (!g_cos || !tmpl_feat || !prop_feat || !g_tmpl_feat || !g_prop_feat) return DMM_ERR_INVALID_ARGUMENT;
  kp.g = g_cos; kp.cos_fwd = cos_fwd; kp.gq = g_tmpl_feat; kp.gk = g_prop_feat;
  const size_t smem = ((size_t)2 * kRowTile * P + 3 * P + 3 * kRowTile) * sizeof(float);
  if (smem > 48 * 1024) return DMM_ERR_UNSUPPORTED_SHAPE;
  cosine_bwd_kernel<<<B, kThreads, smem, (cudaStream_t)stream>>>(kp);
  return check_launch();
}
