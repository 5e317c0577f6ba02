// K4 -- assignment apply: full_outmask = Bmat @ proposal masks, forward and backward.
//
// Reference: dmm/modules/match_model.py:144 (torch.mm of the [O,P] assignment with the [P,HW] soft masks) and the
// valid-row scatter dmm/modules/dmm_model.py:133-135 (torch.mm with a 0/1 [F,O] matrix), fused through row_map.
//
// HBM-bound streaming kernel.  The assignment is sparse by construction (is_test: the row maxima only;
// training: entries > 0.01), so instead of a skinny SGEMM that reads all P masks (22.9 MB at P=50) the kernel
// builds the per-row non-zero list in shared memory and streams only those rows: <= O rows read + O rows written.
#include "common.cuh"

namespace dmm {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxIn = 128;   // max input rows (proposals, or templates in the transposed use)
constexpr int kMaxOut = 128;  // max output rows

struct ApplyParams {
  const float* coef;            // [B][coef_rows][coef_cols]
  long long coef_bs;
  int coef_rs, coef_cs;         // element strides of (out row, in row) inside one problem's matrix
  const float* in;              // [B][*][HW]
  const float* const* in_ptrs;  // optional per-problem base pointers; overrides in/in_bs
  long long in_bs;
  float* out;                   // [B][n_out_rows][HW]
  long long out_bs;
  const int* in_map;            // logical in row -> physical in row   (per problem [n_in_max]) or NULL
  const int* out_map;           // logical out row -> physical out row (per problem [n_out_max]) or NULL
  const int* n_in_arr;          // per-problem logical in rows  (NULL = n_in_max)
  const int* n_out_arr;         // per-problem logical out rows (NULL = n_out_max)
  int n_in_max, n_out_max;      // logical maxima (= stride of the maps)
  int out_rows_phys;            // physical output rows (zero-filled when not produced and zero_fill)
  int zero_fill;
  int B, HW, S;
  int per;                      // work units (float4 / pixels) per slab; 0 = balanced split of the row over S slabs
};

template <bool VEC>
__global__ void __launch_bounds__(kThreads) assign_apply_kernel(const ApplyParams p) {
  __shared__ float nz_val[kMaxOut][kMaxIn / 4];   // compacted coefficients (up to kMaxIn/4 per row kept here...)
  __shared__ unsigned char nz_idx[kMaxOut][kMaxIn / 4];
  __shared__ int nz_cnt[kMaxOut];
  __shared__ int inv[kMaxOut];                    // physical out row -> logical out row or -1
  __shared__ int overflow;
  const int b = blockIdx.y, s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_in = p.n_in_arr ? clampi(p.n_in_arr[b], 0, p.n_in_max) : p.n_in_max;
  const int n_out = p.n_out_arr ? clampi(p.n_out_arr[b], 0, p.n_out_max) : p.n_out_max;
  const float* coef = p.coef + (long long)b * p.coef_bs;
  if (tid == 0) overflow = 0;
  for (int i = tid; i < p.out_rows_phys; i += kThreads) inv[i] = -1;
  __syncthreads();
  for (int o = tid; o < n_out; o += kThreads) {
    const int f = p.out_map ? p.out_map[(long long)b * p.n_out_max + o] : o;
    if (f >= 0 && f < p.out_rows_phys) inv[f] = o;
  }
  constexpr int kCap = kMaxIn / 4;
  for (int o = warp; o < n_out; o += kWarps) {   // ballot-compact the non-zeros of row o
    int cnt = 0;
    for (int c0 = 0; c0 < n_in; c0 += 32) {
      const int c = c0 + lane;
      const float v = c < n_in ? coef[(long long)o * p.coef_rs + (long long)c * p.coef_cs] : 0.f;
      const unsigned mk = __ballot_sync(0xffffffffu, v != 0.f);
      const int pos = cnt + __popc(mk & ((1u << lane) - 1u));
      if (v != 0.f) {
        if (pos < kCap) { nz_val[o][pos] = v; nz_idx[o][pos] = (unsigned char)c; }
        else overflow = 1;
      }
      cnt += __popc(mk);
    }
    if (lane == 0) nz_cnt[o] = cnt;
  }
  __syncthreads();
  const bool dense = overflow != 0;               // rare: a row with more than kCap non-zeros -> read coefficients directly

  const float* inb = p.in_ptrs ? p.in_ptrs[b] : p.in + (long long)b * p.in_bs;
  float* outb = p.out + (long long)b * p.out_bs;
  const int* imap = p.in_map ? p.in_map + (long long)b * p.n_in_max : nullptr;
  const int step = VEC ? 4 : 1;
  const int nq = (p.HW + step - 1) / step;
  const int per = p.per > 0 ? p.per : (nq + p.S - 1) / p.S;
  const int q0 = s * per, q1 = min(q0 + per, nq);
  for (int f = 0; f < p.out_rows_phys; ++f) {
    const int o = inv[f];
    if (o < 0 && !p.zero_fill) continue;
    float* orow = outb + (long long)f * p.HW;
    const int cnt = o < 0 ? 0 : (dense ? n_in : nz_cnt[o]);
    if (VEC && !dense && cnt == 1) {
      // the common inference case (one selected proposal per template): a scaled streaming copy, 4 loads in flight
      const float v = nz_val[o][0];
      const int c = nz_idx[o][0];
      const float* src = inb + (long long)(imap ? imap[c] : c) * p.HW;
      int q = q0 + tid;
      for (; q + 3 * kThreads < q1; q += 4 * kThreads) {
        float4 m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = ld_stream_f4(src + 4 * (q + k * kThreads));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          st_stream_f4(orow + 4 * (q + k * kThreads), make_float4(v * m[k].x, v * m[k].y, v * m[k].z, v * m[k].w));
      }
      for (; q < q1; q += kThreads) {
        const float4 m = ld_stream_f4(src + 4 * q);
        st_stream_f4(orow + 4 * q, make_float4(v * m.x, v * m.y, v * m.z, v * m.w));
      }
      continue;
    }
    if (VEC && cnt == 0) {
      int q = q0 + tid;
      for (; q < q1; q += kThreads) st_stream_f4(orow + 4 * q, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    for (int q = q0 + tid; q < q1; q += kThreads) {
      if (VEC) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = 0; e < cnt; ++e) {
          const int c = dense ? e : nz_idx[o][e];
          const float v = dense ? coef[(long long)o * p.coef_rs + (long long)c * p.coef_cs] : nz_val[o][e];
          if (dense && v == 0.f) continue;
          const int pr = imap ? imap[c] : c;
          const float4 m = ld_stream_f4(inb + (long long)pr * p.HW + 4 * q);
          acc.x = fmaf(v, m.x, acc.x); acc.y = fmaf(v, m.y, acc.y);
          acc.z = fmaf(v, m.z, acc.z); acc.w = fmaf(v, m.w, acc.w);
        }
        st_stream_f4(orow + 4 * q, acc);
      } else {
        float acc = 0.f;
        for (int e = 0; e < cnt; ++e) {
          const int c = dense ? e : nz_idx[o][e];
          const float v = dense ? coef[(long long)o * p.coef_rs + (long long)c * p.coef_cs] : nz_val[o][e];
          if (dense && v == 0.f) continue;
          const int pr = imap ? imap[c] : c;
          acc = fmaf(v, inb[(long long)pr * p.HW + q], acc);
        }
        orow[q] = acc;
      }
    }
  }
}

// ---- backward w.r.t. the assignment: g_B[o,p] = <g_out[row(o)], prop[p]> for the selected entries -----------
struct ApplyBwdParams {
  const float* gout; long long gout_bs;
  const float* prop; long long prop_bs;
  const float* const* prop_ptrs;
  const float* logic;           // [B][O][MS]
  const int* row_map;
  const int* n_prop; const int* n_tmpl;
  int B, P, O, MS, HW, S;
  float* partial;               // [B][S][O*MS]
  float* gB;                    // [B][O][MS]
};

template <bool VEC>
__global__ void __launch_bounds__(kThreads) assign_apply_bwd_partial_kernel(const ApplyBwdParams p) {
  const int b = blockIdx.y, s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  const float* logic = p.logic + (long long)b * p.O * p.MS;
  const float* goutb = p.gout + (long long)b * p.gout_bs;
  const float* propb = p.prop_ptrs ? p.prop_ptrs[b] : p.prop + (long long)b * p.prop_bs;
  float* part = p.partial + ((long long)b * p.S + s) * p.O * p.MS;
  const int step = VEC ? 4 : 1;
  const int nq = (p.HW + step - 1) / step;
  const int per = (nq + p.S - 1) / p.S;
  const int q0 = s * per, q1 = min(q0 + per, nq);
  for (int e = warp; e < nt * np; e += kWarps) {  // one warp per selected entry
    const int o = e / np, c = e - o * np;
    if (logic[o * p.MS + c] == 0.f) continue;     // warp-uniform
    const int f = p.row_map ? p.row_map[(long long)b * p.O + o] : o;
    const float* g = goutb + (long long)f * p.HW;
    const float* m = propb + (long long)c * p.HW;
    float acc = 0.f;
    for (int q = q0 + lane; q < q1; q += 32) {
      if (VEC) {
        const float4 a = *reinterpret_cast<const float4*>(g + 4 * q);
        const float4 d = *reinterpret_cast<const float4*>(m + 4 * q);
        acc = fmaf(a.x, d.x, acc); acc = fmaf(a.y, d.y, acc); acc = fmaf(a.z, d.z, acc); acc = fmaf(a.w, d.w, acc);
      } else {
        acc = fmaf(g[q], m[q], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) part[o * p.MS + c] = acc;
  }
}

__global__ void __launch_bounds__(256) assign_apply_bwd_reduce_kernel(const ApplyBwdParams p) {
  const long long total = (long long)p.B * p.O * p.MS;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.MS);
    const int o = (int)((i / p.MS) % p.O);
    const int b = (int)(i / ((long long)p.MS * p.O));
    const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
    const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
    float r = 0.f;
    if (o < nt && c < np && p.logic[i] != 0.f) {
      const float* part = p.partial + (long long)b * p.S * p.O * p.MS + o * p.MS + c;
      for (int s = 0; s < p.S; ++s) r += part[(long long)s * p.O * p.MS];  // fixed order: deterministic
    }
    p.gB[i] = r;
  }
}

int pick_slabs(int B, int HW) {
  long long want = (8LL * kNumSMs + B - 1) / B;  // ~8 CTAs per SM in flight
  long long maxs = HW / 2048 > 0 ? HW / 2048 : 1;
  long long S = want < 1 ? 1 : (want > maxs ? maxs : want);
  return (int)(S > 65535 ? 65535 : S);
}

// Forward apply: slabs of exactly 4096 pixels (1024 float4 = 4 per thread) whenever the row is long enough, so that every
// thread of every CTA does the same number of 4-deep load batches (the balanced split above left a ragged second batch:
// 1.5 batches per row at B=64; measured 5.2 -> 6.6 TB/s at B=512, the read+write copy peak of the part).
int pick_slabs_fwd(int B, int HW) {
  const long long nq = ((long long)HW + 3) / 4;
  if (nq < 2048) return pick_slabs(B, HW);
  long long S = (nq + 1023) / 1024;
  return (int)(S > 65535 ? 65535 : S);
}

inline bool aligned16(const void* q) { return ((uintptr_t)q & 15u) == 0; }

}  // namespace
}  // namespace dmm

using namespace dmm;

static int run_apply(const float* Bmat, const float* prop, const float* const* prop_ptrs, int ptrs_aligned16,
                     long long prop_bstride, int B, int P, int O, int MS, int HW, const int* n_prop, const int* n_tmpl,
                     const int* row_map, int O_out, int zero_fill, float* out, long long out_bstride, void* stream) {
  if (B < 0 || P < 0 || O < 0 || HW < 0 || MS < P || O_out < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || O_out == 0 || HW == 0) return DMM_OK;
  if (!out) return DMM_ERR_INVALID_ARGUMENT;
  if ((O > 0 && P > 0) && (!Bmat || (!prop && !prop_ptrs))) return DMM_ERR_INVALID_ARGUMENT;
  if (P > kMaxIn || O > kMaxOut || O_out > kMaxOut || B > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;
  ApplyParams kp;
  kp.coef = Bmat; kp.coef_bs = (long long)O * MS; kp.coef_rs = MS; kp.coef_cs = 1;
  kp.in = prop; kp.in_ptrs = prop_ptrs; kp.in_bs = prop_bstride; kp.out = out; kp.out_bs = out_bstride;
  kp.in_map = nullptr; kp.out_map = row_map; kp.n_in_arr = n_prop; kp.n_out_arr = n_tmpl;
  kp.n_in_max = P; kp.n_out_max = O; kp.out_rows_phys = O_out; kp.zero_fill = zero_fill;
  const bool vec = HW % 4 == 0 && (prop_ptrs ? ptrs_aligned16 != 0 : (aligned16(prop) && prop_bstride % 4 == 0)) &&
                   aligned16(out) && out_bstride % 4 == 0;
  kp.B = B; kp.HW = HW; kp.per = 0;
  kp.S = vec ? pick_slabs_fwd(B, HW) : pick_slabs(B, HW);
  if (vec && kp.S == (int)((((long long)HW + 3) / 4 + 1023) / 1024) && (long long)HW >= 8192) kp.per = 1024;
  dim3 grid(kp.S, B);
  if (vec) assign_apply_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(kp);
  else assign_apply_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(kp);
  return check_launch();
}

extern "C" int dmm_assign_apply(const float* Bmat, const float* prop, long long prop_bstride, int B, int P, int O,
                                int MS, int HW, const int* n_prop, const int* n_tmpl, const int* row_map, int O_out,
                                int zero_fill, float* out, long long out_bstride, void* stream) {
  return run_apply(Bmat, prop, nullptr, 0, prop_bstride, B, P, O, MS, HW, n_prop, n_tmpl, row_map, O_out, zero_fill, out,
                   out_bstride, stream);
}

extern "C" int dmm_assign_apply_ptrs(const float* Bmat, const float* const* prop_ptrs, int ptrs_aligned16, int B, int P,
                                     int O, int MS, int HW, const int* n_prop, const int* n_tmpl, const int* row_map,
                                     int O_out, int zero_fill, float* out, long long out_bstride, void* stream) {
  if (!prop_ptrs) return DMM_ERR_INVALID_ARGUMENT;
  return run_apply(Bmat, nullptr, prop_ptrs, ptrs_aligned16, 0, B, P, O, MS, HW, n_prop, n_tmpl, row_map, O_out,
                   zero_fill, out, out_bstride, stream);
}

extern "C" size_t dmm_assign_apply_bwd_workspace_bytes(int B, int P, int O, int HW) {
  if (B <= 0 || O <= 0 || HW <= 0) return 256;
  const int MS = P > O + 1 ? P : O + 1;
  return align_up((size_t)B * pick_slabs(B, HW) * O * MS * sizeof(float), 256);
}

static int run_apply_bwd(const float* g_out, long long gout_bstride, const float* prop, const float* const* prop_ptrs,
                         int ptrs_aligned16, long long prop_bstride, const float* Bmat, const float* logic, int B, int P,
                         int O, int MS, int HW, const int* n_prop, const int* n_tmpl, const int* row_map, float* g_Bmat,
                         float* g_prop, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || P < 0 || O < 0 || HW < 0 || MS < P) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || O == 0 || P == 0) return DMM_OK;
  if (!g_out || (!prop && !prop_ptrs) || !logic) return DMM_ERR_INVALID_ARGUMENT;
  if (P > kMaxIn || O > kMaxOut || B > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = HW % 4 == 0 && (prop_ptrs ? ptrs_aligned16 != 0 : (aligned16(prop) && prop_bstride % 4 == 0)) &&
                   aligned16(g_out) && gout_bstride % 4 == 0;
  if (g_Bmat) {
    ApplyBwdParams kp;
    kp.gout = g_out; kp.gout_bs = gout_bstride; kp.prop = prop; kp.prop_ptrs = prop_ptrs; kp.prop_bs = prop_bstride;
    kp.logic = logic;
    kp.row_map = row_map; kp.n_prop = n_prop; kp.n_tmpl = n_tmpl;
    kp.B = B; kp.P = P; kp.O = O; kp.MS = MS; kp.HW = HW; kp.S = pick_slabs(B, HW > 0 ? HW : 1);
    if (!workspace || workspace_bytes < (size_t)B * kp.S * O * MS * sizeof(float)) return DMM_ERR_WORKSPACE_TOO_SMALL;
    kp.partial = (float*)workspace; kp.gB = g_Bmat;
    dim3 grid(kp.S, B);
    if (vec) assign_apply_bwd_partial_kernel<true><<<grid, kThreads, 0, st>>>(kp);
    else assign_apply_bwd_partial_kernel<false><<<grid, kThreads, 0, st>>>(kp);
    int rc = check_launch();
    if (rc) return rc;
    const long long total = (long long)B * O * MS;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    assign_apply_bwd_reduce_kernel<<<blocks, 256, 0, st>>>(kp);
    rc = check_launch();
    if (rc) return rc;
  }
  if (g_prop) {  // g_prop[p] = sum_o Bmat[o,p] * g_out[row(o)]: the forward kernel with the transposed matrix
    if (!Bmat) return DMM_ERR_INVALID_ARGUMENT;
    ApplyParams kp;
    kp.coef = Bmat; kp.coef_bs = (long long)O * MS; kp.coef_rs = 1; kp.coef_cs = MS;
    kp.in = g_out; kp.in_ptrs = nullptr; kp.in_bs = gout_bstride; kp.out = g_prop; kp.out_bs = (long long)P * HW;  // g_prop is dense [B][P][HW]
    kp.in_map = row_map; kp.out_map = nullptr; kp.n_in_arr = n_tmpl; kp.n_out_arr = n_prop;
    kp.n_in_max = O; kp.n_out_max = P; kp.out_rows_phys = P; kp.zero_fill = 1;
    kp.B = B; kp.HW = HW; kp.S = pick_slabs(B, HW > 0 ? HW : 1); kp.per = 0;
    const bool vec2 = vec && aligned16(g_prop);
    dim3 grid(kp.S, B);
    if (vec2) assign_apply_kernel<true><<<grid, kThreads, 0, st>>>(kp);
    else assign_apply_kernel<false><<<grid, kThreads, 0, st>>>(kp);
    return check_launch();
  }
  return DMM_OK;
}

extern "C" int dmm_assign_apply_bwd(const float* g_out, long long gout_bstride, const float* prop,
                                    long long prop_bstride, const float* Bmat, const float* logic, int B, int P,
                                    int O, int MS, int HW, const int* n_prop, const int* n_tmpl, const int* row_map,
                                    float* g_Bmat, float* g_prop, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  return run_apply_bwd(g_out, gout_bstride, prop, nullptr, 0, prop_bstride, Bmat, logic, B, P, O, MS, HW, n_prop, n_tmpl,
                       row_map, g_Bmat, g_prop, workspace, workspace_bytes, stream);
}

/* g_prop is not offered here: per-video proposal tensors come from the (non-differentiable) proposal generator. */
extern "C" int dmm_assign_apply_bwd_ptrs(const float* g_out, long long gout_bstride, const float* const* prop_ptrs,
                                         int ptrs_aligned16, const float* Bmat, const float* logic, int B, int P, int O,
                                         int MS, int HW, const int* n_prop, const int* n_tmpl, const int* row_map,
                                         float* g_Bmat, void* workspace, size_t workspace_bytes, void* stream) {
  if (!prop_ptrs) return DMM_ERR_INVALID_ARGUMENT;
  return run_apply_bwd(g_out, gout_bstride, nullptr, prop_ptrs, ptrs_aligned16, 0, Bmat, logic, B, P, O, MS, HW, n_prop,
                       n_tmpl, row_map, g_Bmat, nullptr, workspace, workspace_bytes, stream);
}
