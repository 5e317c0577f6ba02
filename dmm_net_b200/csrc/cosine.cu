// K2 -- pairwise cosine similarity between template and proposal features, forward and backward.
//
// Reference: dmm/utils/match_helper.py:51-64 (F.cosine_similarity over D on expanded [O,D,P] views, eps 1e-8)
// averaged over the template-feature sets (dmm/modules/match_model.py:72-76; one set in practice,
// dmm/modules/dmm_model.py:44).  ATen normalises each vector by max(||v||, eps) first, multiplies, then sums.
//
// 0.5 MFLOP and 123 KB per match next to 27.5 MB of masks: this kernel is deliberately plain fp32 FFMA
// (a TF32 tensor-core contraction would break the 1e-4 parity bar for nothing).  One CTA per problem; the
// normalised template rows of a tile sit in shared memory, each warp owns proposals.
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
  float* gq;        // bwd: [B][T][O][D]
  float* gk;        // bwd: [B][P][D]
};

__device__ __forceinline__ float row_norm(const float* v, int D, int lane) {
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(v[d], v[d], s);
  return __fsqrt_rn(warp_sum(s));
}

__global__ void __launch_bounds__(kThreads) cosine_fwd_kernel(const CosParams p) {
  extern __shared__ float qs[];  // [rt][D] normalised template rows
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  const int D = p.D;
  int rt = kSmemFloats / (D > 0 ? D : 1);
  rt = rt > kRowTile ? kRowTile : (rt < 1 ? 1 : rt);
  float* cosb = p.cos + (long long)b * p.O * p.P;
  for (int i = tid; i < p.O * p.P; i += kThreads) cosb[i] = 0.f;
  __syncthreads();
  const float* kb = p.k + (long long)b * p.P * D;
  for (int t = 0; t < p.T; ++t) {
    const float* qb = p.q + ((long long)b * p.T + t) * p.O * D;
    for (int o0 = 0; o0 < nt; o0 += rt) {
      const int ro = min(rt, nt - o0);
      __syncthreads();
      for (int r = warp; r < ro; r += kWarps) {
        const float* v = qb + (long long)(o0 + r) * D;
        const float nq = fmaxf(row_norm(v, D, lane), p.eps);
        for (int d = lane; d < D; d += 32) qs[r * D + d] = __fdiv_rn(v[d], nq);
      }
      __syncthreads();
      for (int c = warp; c < np; c += kWarps) {
        const float* kv = kb + (long long)c * D;
        const float nk = fmaxf(row_norm(kv, D, lane), p.eps);
        float acc[kRowTile];
#pragma unroll
        for (int r = 0; r < kRowTile; ++r) acc[r] = 0.f;
        for (int d = lane; d < D; d += 32) {
          const float kn = __fdiv_rn(kv[d], nk);
#pragma unroll
          for (int r = 0; r < kRowTile; ++r)
            if (r < ro) acc[r] = fmaf(qs[r * D + d], kn, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kRowTile; ++r)
          if (r < ro) {
            const float s = warp_sum(acc[r]);
            // this warp is the only writer of column c; the first template set stores (a read-modify-write here is a
            // dependent ~0.7 us global round trip per entry: it made this kernel 60 us at B=64), later sets accumulate
            if (lane == 0) {
              if (t == 0) cosb[(o0 + r) * p.P + c] = s;
              else cosb[(o0 + r) * p.P + c] += s;
            }
          }
      }
    }
  }
  __syncthreads();
  if (p.T > 1) {
    const float tf = (float)p.T;
    for (int i = tid; i < p.O * p.P; i += kThreads) cosb[i] = __fdiv_rn(cosb[i], tf);  // feature_sim /= T
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
      // phase 1: cos_t of the tile (one warp per (row, proposal) pair)
      for (int i = warp; i < ro * np; i += kWarps) {
        const int r = i / np, c = i - r * np;
        const float* qv = qb + (long long)(o0 + r) * D;
        const float* kv = kb + (long long)c * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(qv[d] / nq[r], kv[d] / nk[c], s);
        s = warp_sum(s);
        if (lane == 0) ct[r * P + c] = s;
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
        for (int c = 0; c < np; ++c) {
          const float kraw = kb[(long long)c * D + d];
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
          gkb[(long long)c * D + d] += (bq - sk[c] * unit) / nk[c];
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

static int fill_params(CosParams& kp, const float* q, const float* k, int B, int T, int P, int O, int D,
                       const int* n_prop, const int* n_tmpl, float eps) {
  if (B < 0 || T < 1 || P < 0 || O < 0 || D < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (D > kSmemFloats) return DMM_ERR_UNSUPPORTED_SHAPE;
  kp.q = q; kp.k = k; kp.B = B; kp.T = T; kp.P = P; kp.O = O; kp.D = D;
  kp.n_prop = n_prop; kp.n_tmpl = n_tmpl; kp.eps = eps;
  kp.cos = nullptr; kp.g = nullptr; kp.gq = nullptr; kp.gk = nullptr;
  return DMM_OK;
}

extern "C" int dmm_cosine_pairwise(const float* tmpl_feat, const float* prop_feat, int B, int T, int P, int O, int D,
                                   const int* n_prop, const int* n_tmpl, float eps, float* cos, void* stream) {
  CosParams kp;
  int rc = fill_params(kp, tmpl_feat, prop_feat, B, T, P, O, D, n_prop, n_tmpl, eps);
  if (rc) return rc;
  if (B == 0 || P == 0 || O == 0) return DMM_OK;
  if (!tmpl_feat || !prop_feat || !cos) return DMM_ERR_INVALID_ARGUMENT;
  kp.cos = cos;
  int rt = kSmemFloats / (D > 0 ? D : 1);
  rt = rt > kRowTile ? kRowTile : (rt < 1 ? 1 : rt);
  const size_t smem = (size_t)rt * (D > 0 ? D : 1) * sizeof(float);
  cosine_fwd_kernel<<<B, kThreads, smem, (cudaStream_t)stream>>>(kp);
  return check_launch();
}

extern "C" int dmm_cosine_pairwise_bwd(const float* g_cos, const float* tmpl_feat, const float* prop_feat, int B,
                                       int T, int P, int O, int D, const int* n_prop, const int* n_tmpl, float eps,
                                       float* g_tmpl_feat, float* g_prop_feat, void* stream) {
  CosParams kp;
  int rc = fill_params(kp, tmpl_feat, prop_feat, B, T, P, O, D, n_prop, n_tmpl, eps);
  if (rc) return rc;
  if (B == 0 || P == 0 || O == 0 || D == 0) return DMM_OK;
  if (!g_cos || !tmpl_feat || !prop_feat || !g_tmpl_feat || !g_prop_feat) return DMM_ERR_INVALID_ARGUMENT;
  kp.g = g_cos; kp.gq = g_tmpl_feat; kp.gk = g_prop_feat;
  const size_t smem = ((size_t)2 * kRowTile * P + 3 * P + 3 * kRowTile) * sizeof(float);
  if (smem > 48 * 1024) return DMM_ERR_UNSUPPORTED_SHAPE;
  cosine_bwd_kernel<<<B, kThreads, smem, (cudaStream_t)stream>>>(kp);
  return check_launch();
}
