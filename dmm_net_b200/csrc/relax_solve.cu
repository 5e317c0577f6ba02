// K3 -- relaxed matching solver + assignment head, forward and backward.
//
// Reference: dmm/modules/submodules/relax_match.py:36-105 (greedy init :45-55, gradient step :68-71,
// Dykstra sweep :73-87 with project_col :21-34 and project_row :9-19, inner exit :88-89, outer exit :96-98)
// and the [O x m] part of dmm/modules/match_model.py:93-148 (pad rule :109-113, mean of X_list :121,
// logic mask :125-130, match_score :146, det_score :147).
//
// One CTA of 4 warps per problem (round-1 profile: one warp per problem was a 478 us single-warp latency chain).
// Warp w owns a contiguous block of rows, lane l owns columns l, l+32, ...: every thread keeps its NRW x CPL
// elements of X, the three Dykstra increments, C and the running sum of iterates in registers for the whole solve.
// Row sums are intra-warp xor-butterflies (bit-identical in every lane); column sums are per-warp partials
// combined through shared memory in a fixed order (bit-identical in every thread), so both data-dependent exits
// are block-uniform and evaluated on the device with the reference's exact-equality tests -- no host syncs.
// Arithmetic mirrors the reference's fp32 op order with explicit *_rn intrinsics (no FMA contraction).
#include "common.cuh"

namespace dmm {
namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kMaxRows = 16;
constexpr int kMaxCols = 128;

struct SolveParams {
  const float* mat;
  const float* score;
  int B, P, O, MS;
  const int* n_prop;
  const int* n_tmpl;
  int max_iter, proj_iter;
  float lr;
  int negate, pad_rule, is_test;
  float *R, *Xf, *Bm, *logic, *ms, *ds;
  int* n_list;
  float* xlist;
  float* cost;
  uint32_t* saved_bits;  // [B][max_iter*proj_iter][128]
  int* saved_sweeps;     // [B][max_iter]
};

__device__ __forceinline__ void problem_dims(const int* n_prop, const int* n_tmpl, int b, int P, int O,
                                             int pad_rule, int& n, int& np, int& m) {
  np = n_prop ? clampi(n_prop[b], 0, P) : P;
  n = n_tmpl ? clampi(n_tmpl[b], 0, O) : O;
  m = (pad_rule && np <= n) ? n + 1 : np;
}

// fixed-order combine of the four per-warp partials: same bits in every thread
__device__ __forceinline__ float sum4(const float* s) { return __fadd_rn(__fadd_rn(__fadd_rn(s[0], s[1]), s[2]), s[3]); }

// small register tiles ask for 6 CTAs/SM (<= 85 registers, no spills): 888 problems in flight per wave instead of 592
template <int NRW, int CPL>
__global__ void __launch_bounds__(kThreads, (NRW * CPL <= 6) ? 6 : 1) relax_solve_kernel(const SolveParams p) {
  constexpr int W = CPL * 32;
  __shared__ float s_col[kWarps][W];   // per-warp column partials (sums / minima)
  __shared__ int s_idx[kWarps][W];
  __shared__ float s_red[2][kWarps];   // per-warp scalars (max / cost), double-buffered
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int r0 = warp * NRW;           // first row of this warp

  int n, np, m;
  problem_dims(p.n_prop, p.n_tmpl, b, p.P, p.O, p.pad_rule, n, np, m);
  const float nf = (float)n, mf = (float)m;
  bool cv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) cv[j] = lane + 32 * j < m;

  const long long obase = (long long)b * p.O * p.MS;
  auto write_mat = [&](float* dst, int r, int j, float v) {
    const int col = lane + 32 * j;
    if (dst && r < p.O && col < p.MS) dst[obase + (long long)r * p.MS + col] = v;
  };
  if (n == 0 || m == 0) {  // nothing to match (block-uniform): define every output as zero
#pragma unroll
    for (int i = 0; i < NRW; ++i) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        write_mat(p.R, r0 + i, j, 0.f); write_mat(p.Xf, r0 + i, j, 0.f);
        write_mat(p.Bm, r0 + i, j, 0.f); write_mat(p.logic, r0 + i, j, 0.f);
      }
      if (lane == 0 && r0 + i < p.O) {
        if (p.ms) p.ms[(long long)b * p.O + r0 + i] = 0.f;
        if (p.ds) p.ds[(long long)b * p.O + r0 + i] = 0.f;
      }
    }
    if (threadIdx.x == 0 && p.n_list) p.n_list[b] = 0;
    return;
  }

  // ---- load the cost; C.max() and the per-column best row (relax_match.py:45-51) ------------------------
  const float* mat = p.mat + (long long)b * p.O * p.P;
  float c[NRW][CPL];
  float lmax = -INFINITY;
  float cmin[CPL];
  int cidx[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { cmin[j] = INFINITY; cidx[j] = 0x7fffffff; }
#pragma unroll
  for (int i = 0; i < NRW; ++i) {
    const int r = r0 + i;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      float v = 0.f;
      if (r < n && col < np) {
        v = mat[r * p.P + col];
        if (p.negate) v = -v;
      }
      c[i][j] = v;
      if (r < n && cv[j]) {
        lmax = fmaxf(lmax, v);
        if (v < cmin[j]) { cmin[j] = v; cidx[j] = r; }  // rows ascending: first minimum wins
      }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) s_red[0][warp] = lmax;
#pragma unroll
  for (int j = 0; j < CPL; ++j) { s_col[warp][lane + 32 * j] = cmin[j]; s_idx[warp][lane + 32 * j] = cidx[j]; }
  __syncthreads();
  const float fill = fmaxf(fmaxf(s_red[0][0], s_red[0][1]), fmaxf(s_red[0][2], s_red[0][3]));
  int best[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    float bv = INFINITY;
    best[j] = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {  // warps hold ascending row blocks: strict < keeps the first minimum
      const float v = s_col[w][lane + 32 * j];
      if (v < bv) { bv = v; best[j] = s_idx[w][lane + 32 * j]; }
    }
  }
  __syncthreads();  // s_col / s_red are reused below

  // ---- greedy start (relax_match.py:52-55) -------------------------------------------------------------
  float x[NRW][CPL], q0[NRW][CPL], q1[NRW][CPL], q2[NRW][CPL], racc[NRW][CPL];
#pragma unroll
  for (int i = 0; i < NRW; ++i) {
    const int r = r0 + i;
    float bv = INFINITY;
    int bi = 0x7fffffff;
    if (r < n) {
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (cv[j]) {
          const float kept = best[j] == r ? c[i][j] : fill;
          if (kept < bv) { bv = kept; bi = lane + 32 * j; }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // lexicographic (value, column) minimum == first minimum
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      x[i][j] = (r < n && lane + 32 * j == bi) ? 1.f : 0.f;
      q0[i][j] = q1[i][j] = q2[i][j] = 0.f;
      racc[i][j] = x[i][j];  // sum(X_list) starts as 0 + X0 == X0
    }
  }
  int L = 1;
  float* xl = p.xlist ? p.xlist + (long long)b * (p.max_iter + 1) * p.O * p.MS : nullptr;
  float* costv = p.cost ? p.cost + (long long)b * (p.max_iter + 1) : nullptr;
  auto record_iterate = [&](int slot) {
    if (!xl) return;
    float* dst = xl + (long long)slot * p.O * p.MS;
#pragma unroll
    for (int i = 0; i < NRW; ++i)
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int r = r0 + i, col = lane + 32 * j;
        if (r < p.O && col < p.MS) dst[r * p.MS + col] = (r < n && cv[j]) ? x[i][j] : 0.f;
      }
  };
  record_iterate(0);
  if (costv && threadIdx.x == 0) costv[0] = 0.f;

  uint32_t* sbits = p.saved_bits ? p.saved_bits + (long long)b * p.max_iter * p.proj_iter * kThreads : nullptr;
  int* ssweeps = p.saved_sweeps ? p.saved_sweeps + (long long)b * p.max_iter : nullptr;

  float cost_prev = 0.f;
  for (int it = 0; it < p.max_iter; ++it) {
    // ---- gradient step, cost, record (relax_match.py:69-71) ------------------------------------------
    float c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NRW; ++i)
      if (r0 + i < n) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          x[i][j] = __fsub_rn(x[i][j], __fmul_rn(p.lr, c[i][j]));
          const float xc = __fmul_rn(x[i][j], c[i][j]);
          c2 = __fadd_rn(c2, __fmul_rn(xc, xc));
          racc[i][j] = __fadd_rn(racc[i][j], x[i][j]);
        }
      }
    c2 = warp_sum(c2);
    if (lane == 0) s_red[it & 1][warp] = c2;
    record_iterate(L);
    ++L;
    __syncthreads();
    const float cost_cur = __fsqrt_rn(sum4(s_red[it & 1]));
    if (costv && threadIdx.x == 0) costv[L - 1] = cost_cur;

    // ---- Dykstra sweeps (relax_match.py:73-89) ----------------------------------------------------------
    int sweeps = 0;
    for (int js = 0; js < p.proj_iter; ++js) {
      float xs[NRW][CPL];
      float cs[CPL];
      uint32_t sb = 0u;
#pragma unroll
      for (int j = 0; j < CPL; ++j) cs[j] = 0.f;
#pragma unroll
      for (int i = 0; i < NRW; ++i)
        if (r0 + i < n) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            xs[i][j] = x[i][j];
            const float a = __fadd_rn(x[i][j], q0[i][j]);
            const float y = fmaxf(a, 0.f);                      // {X >= 0}
            q0[i][j] = __fsub_rn(a, y);
            const float bb = __fadd_rn(y, q1[i][j]);
            x[i][j] = bb;
            cs[j] = __fadd_rn(cs[j], bb);                       // column partial over this warp's rows, in order
            if (a > 0.f) sb |= 1u << (i * CPL + j);
          }
        }
#pragma unroll
      for (int j = 0; j < CPL; ++j) s_col[warp][lane + 32 * j] = cs[j];
      __syncthreads();
      bool keep[CPL];
      float tcol[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int col = lane + 32 * j;
        const float tot = __fadd_rn(__fadd_rn(__fadd_rn(s_col[0][col], s_col[1][col]), s_col[2][col]), s_col[3][col]);
        keep[j] = tot <= 1.f;                                   // {col sums <= 1}: only violators move
        tcol[j] = __fdiv_rn(__fsub_rn(tot, 1.f), nf);
        if (keep[j]) sb |= 1u << (NRW * CPL + j);
      }
      float rs[NRW];
#pragma unroll
      for (int i = 0; i < NRW; ++i) {
        float acc = 0.f;
        if (r0 + i < n) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float bb = x[i][j];
            const float y1 = keep[j] ? bb : __fsub_rn(bb, tcol[j]);
            q1[i][j] = __fsub_rn(bb, y1);
            const float cc = __fadd_rn(y1, q2[i][j]);
            x[i][j] = cc;
            acc = __fadd_rn(acc, cc);
          }
        }
        rs[i] = acc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NRW; ++i) rs[i] = __fadd_rn(rs[i], __shfl_xor_sync(0xffffffffu, rs[i], o));
      }
      int changed = 0;
#pragma unroll
      for (int i = 0; i < NRW; ++i)
        if (r0 + i < n) {
          const float u = __fdiv_rn(__fsub_rn(rs[i], 1.f), mf);  // {row sums == 1}
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float cc = x[i][j];
            const float y2 = cv[j] ? __fsub_rn(cc, u) : 0.f;
            q2[i][j] = __fsub_rn(cc, y2);
            x[i][j] = y2;
            const float d = __fsub_rn(y2, xs[i][j]);
            changed |= __fmul_rn(d, d) != 0.f;                    // ||X - X_start|| == 0  <=>  every square is 0
          }
        }
      if (sbits) sbits[((long long)it * p.proj_iter + js) * kThreads + threadIdx.x] = sb;
      ++sweeps;
      // block-wide OR; also orders this sweep's reads of s_col before the next sweep's writes
      if (!__syncthreads_or(changed)) break;
    }
    if (ssweeps && threadIdx.x == 0) ssweeps[it] = sweeps;
    if (cost_prev == cost_cur) break;                             // relax_match.py:96 (block-uniform)
    cost_prev = cost_cur;
  }

  // ---- head: R = mean(X_list), logic, Bmat, scores (match_model.py:121-147) ----------------------------
  const float Lf = (float)L;
  const float* score = p.score ? p.score + (long long)b * p.P : nullptr;
#pragma unroll
  for (int i = 0; i < NRW; ++i) {
    const int r = r0 + i;
    if (r >= p.O) continue;
    const bool rv = r < n;
    float Rv[CPL], top = -INFINITY;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      Rv[j] = (rv && cv[j]) ? __fdiv_rn(racc[i][j], Lf) : 0.f;
      if (rv && cv[j]) top = fmaxf(top, Rv[j]);
    }
    top = warp_max(top);
    float bestv = -INFINITY;
    double det = 0.0;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      float lg = 0.f, bm = 0.f;
      if (rv && cv[j]) {
        lg = p.is_test ? (Rv[j] == top ? 1.f : 0.f) : (Rv[j] > 0.01f ? 1.f : 0.f);
        bm = __fmul_rn(Rv[j], lg);
        const float simv = -c[i][j];                            // (-cost_matrix), match_model.py:146
        bestv = fmaxf(bestv, __fmul_rn(fminf(fmaxf(Rv[j], 0.f), 1.f), simv));
        if (score && col < np) det += (double)__fmul_rn(score[col], bm);
      }
      write_mat(p.R, r, j, Rv[j]);
      write_mat(p.logic, r, j, lg);
      write_mat(p.Bm, r, j, bm);
      write_mat(p.Xf, r, j, (rv && cv[j]) ? x[i][j] : 0.f);
    }
    bestv = warp_max(bestv);
    det = warp_sum(det);
    if (lane == 0) {
      if (p.ms) p.ms[(long long)b * p.O + r] = rv ? bestv : 0.f;
      if (p.ds) p.ds[(long long)b * p.O + r] = rv ? (float)det : 0.f;
    }
  }
  if (threadIdx.x == 0 && p.n_list) p.n_list[b] = L;
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
struct SolveBwdParams {
  const float *gR, *gXf, *gBm, *gms, *gds;
  const float *mat, *score, *R, *logic;
  const int* n_list;
  const uint32_t* saved_bits;
  const int* saved_sweeps;
  int B, P, O, MS;
  const int* n_prop;
  const int* n_tmpl;
  int max_iter, proj_iter;
  float lr;
  int negate, pad_rule;
  float* g_mat;
  float* g_score;
};

// The sweep is piecewise linear.  With the saved ReLU bits / column-active bits the adjoint recursion is
//   gY2 = gX' - gP2';  gCc = gP2' + gY2 - rowsum(gY2)/m;  gP2 = gCc;  h = gCc - gP1';
//   gBb = gP1' + (keep ? h : h - colsum(h)/n);  gP1 = gBb;  gA = gP0' + relu'(A) * (gBb - gP0');  gP0 = gX = gA
// and per outer step  gC -= lr * (gX + gR/L)  (the recorded iterate carries weight 1/L in R).
template <int NRW, int CPL>
__global__ void __launch_bounds__(kThreads) relax_solve_bwd_kernel(const SolveBwdParams p) {
  constexpr int W = CPL * 32;
  __shared__ float s_col[2][kWarps][W];
  __shared__ float s_gsc[kWarps][W];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int r0 = warp * NRW;
  int n, np, m;
  problem_dims(p.n_prop, p.n_tmpl, b, p.P, p.O, p.pad_rule, n, np, m);
  float* g_mat = p.g_mat + (long long)b * p.O * p.P;
  float* g_score = p.g_score ? p.g_score + (long long)b * p.P : nullptr;
  const int L = p.n_list[b];
  if (n == 0 || m == 0 || L <= 0) {  // block-uniform
    for (int i = threadIdx.x; i < p.O * p.P; i += kThreads) g_mat[i] = 0.f;
    if (g_score) for (int i = threadIdx.x; i < p.P; i += kThreads) g_score[i] = 0.f;
    return;
  }
  const float nf = (float)n, mf = (float)m, Lf = (float)L;
  bool cv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) cv[j] = lane + 32 * j < m;
  const long long obase = (long long)b * p.O * p.MS;
  const float* mat = p.mat + (long long)b * p.O * p.P;
  const float* score = p.score ? p.score + (long long)b * p.P : nullptr;

  // ---- head backward: cotangent of R, the direct path into `mat` (match_score), d/d prop_score ----------
  float gRl[NRW][CPL], gdir[NRW][CPL], gsc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) gsc[j] = 0.f;
#pragma unroll
  for (int i = 0; i < NRW; ++i) {
    const int r = r0 + i;
    const bool rv = r < n;
    const float gds = (rv && p.gds) ? p.gds[(long long)b * p.O + r] : 0.f;
    const float gms = (rv && p.gms) ? p.gms[(long long)b * p.O + r] : 0.f;
    float Rv[CPL], sv[CPL];
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      const long long oi = obase + (long long)r * p.MS + col;
      Rv[j] = (rv && cv[j]) ? p.R[oi] : 0.f;
      sv[j] = (rv && cv[j] && col < np) ? mat[r * p.P + col] : 0.f;
      if (!p.negate) sv[j] = -sv[j];                           // match_score multiplies by -C
      const float val = __fmul_rn(fminf(fmaxf(Rv[j], 0.f), 1.f), sv[j]);
      if (rv && cv[j] && val > best) { best = val; bi = col; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                         // first maximum, like torch.max(dim)
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      float g = 0.f, gd = 0.f;
      if (rv && cv[j]) {
        const long long oi = obase + (long long)r * p.MS + col;
        const float lg = p.logic[oi];
        float gB = p.gBm ? p.gBm[oi] : 0.f;
        const float sc = (score && col < np) ? score[col] : 0.f;
        gB = fmaf(gds, sc, gB);                                // det_score = sum score * Bmat
        gsc[j] = fmaf(gds, Rv[j] * lg, gsc[j]);
        g = (p.gR ? p.gR[oi] : 0.f) + gB * lg;                 // Bmat = R * logic (logic is a constant)
        if (col == bi && gms != 0.f) {                         // match_score = max(clamp(R,0,1) * (-C))
          if (Rv[j] >= 0.f && Rv[j] <= 1.f) g = fmaf(gms, sv[j], g);
          gd = gms * fminf(fmaxf(Rv[j], 0.f), 1.f);
        }
        g = g / Lf;
      }
      gRl[i][j] = g;
      gdir[i][j] = gd;
    }
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) s_gsc[warp][lane + 32 * j] = gsc[j];
  __syncthreads();
  if (g_score && warp == 0) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      if (col < p.P) g_score[col] = col < np ? (s_gsc[0][col] + s_gsc[1][col]) + (s_gsc[2][col] + s_gsc[3][col]) : 0.f;
    }
  }

  // ---- reverse sweep through the solver ---------------------------------------------------------------
  float gx[NRW][CPL], g0[NRW][CPL], g1[NRW][CPL], g2[NRW][CPL], gC[NRW][CPL];
#pragma unroll
  for (int i = 0; i < NRW; ++i)
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int r = r0 + i, col = lane + 32 * j;
      gx[i][j] = (p.gXf && r < n && cv[j]) ? p.gXf[obase + (long long)r * p.MS + col] : 0.f;
      g0[i][j] = g1[i][j] = g2[i][j] = gC[i][j] = 0.f;
    }
  const uint32_t* sbits = p.saved_bits + (long long)b * p.max_iter * p.proj_iter * kThreads;
  const int* ssweeps = p.saved_sweeps + (long long)b * p.max_iter;

  int par = 0;
  for (int it = L - 2; it >= 0; --it) {
    const int sweeps = ssweeps[it];
    for (int js = sweeps - 1; js >= 0; --js) {
      const uint32_t sb = sbits[((long long)it * p.proj_iter + js) * kThreads + threadIdx.x];
      float rs[NRW];
#pragma unroll
      for (int i = 0; i < NRW; ++i) {
        float acc = 0.f;
        if (r0 + i < n) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float gy2 = cv[j] ? gx[i][j] - g2[i][j] : 0.f;
            gx[i][j] = gy2;  // reuse as gY2
            acc += gy2;
          }
        }
        rs[i] = acc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < NRW; ++i) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
      }
      float hs[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) hs[j] = 0.f;
#pragma unroll
      for (int i = 0; i < NRW; ++i)
        if (r0 + i < n) {
          const float u = rs[i] / mf;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float gcc = cv[j] ? g2[i][j] + gx[i][j] - u : 0.f;  // d/dCc
            g2[i][j] = gcc;                                              // Cc = Y1 + P2  -> gP2
            const float h = gcc - g1[i][j];                              // P1' = Bb - Y1
            gx[i][j] = h;                                                // reuse as h
            hs[j] += h;
          }
        }
#pragma unroll
      for (int j = 0; j < CPL; ++j) s_col[par][warp][lane + 32 * j] = hs[j];
      __syncthreads();  // double-buffered partials: one barrier per sweep
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int col = lane + 32 * j;
        hs[j] = ((s_col[par][0][col] + s_col[par][1][col]) + s_col[par][2][col]) + s_col[par][3][col];
      }
      par ^= 1;
#pragma unroll
      for (int i = 0; i < NRW; ++i)
        if (r0 + i < n) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const bool keep = (sb >> (NRW * CPL + j)) & 1u;
            const float h = gx[i][j];
            const float gbb = g1[i][j] + (keep ? h : h - hs[j] / nf);    // Y1 = colproj(Bb)
            g1[i][j] = gbb;                                              // Bb = Y0 + P1 -> gP1
            const float t = gbb - g0[i][j];                              // P0' = A - Y0
            const bool pos = (sb >> (i * CPL + j)) & 1u;
            const float ga = g0[i][j] + (pos ? t : 0.f);                 // Y0 = relu(A)
            g0[i][j] = ga;                                               // A = X + P0 -> gP0
            gx[i][j] = cv[j] ? ga : 0.f;
          }
        }
    }
    // X_g = X_prev - lr*C is both recorded (weight 1/L in R) and fed to the sweeps
#pragma unroll
    for (int i = 0; i < NRW; ++i)
      if (r0 + i < n) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const float g = gx[i][j] + gRl[i][j];
          gC[i][j] = fmaf(-p.lr, g, gC[i][j]);
          gx[i][j] = g;
        }
      }
  }
  // ---- d loss / d mat = direct term -/+ d loss / d C ---------------------------------------------------
#pragma unroll
  for (int i = 0; i < NRW; ++i)
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int r = r0 + i, col = lane + 32 * j;
      if (r < p.O && col < p.P) {
        float g = 0.f;
        if (r < n && col < np) g = p.negate ? (gdir[i][j] - gC[i][j]) : (gC[i][j] - gdir[i][j]);
        g_mat[r * p.P + col] = g;
      }
    }
}

template <typename F>
int dispatch_shape(int O, int MS, F&& f) {
  // per-thread register tile: NRW rows per warp (4 warps) x CPL columns per lane
  const int nrw = (O + kWarps - 1) / kWarps, cpl = (MS + 31) / 32;
  if (O > kMaxRows || MS > kMaxCols || O < 1 || MS < 1) return DMM_ERR_UNSUPPORTED_SHAPE;
#define DMM_CASE(NR, CP) \
  if (nrw <= NR && cpl <= CP) return f(std::integral_constant<int, NR>{}, std::integral_constant<int, CP>{});
  DMM_CASE(1, 1) DMM_CASE(1, 2) DMM_CASE(2, 1) DMM_CASE(2, 2) DMM_CASE(3, 2) DMM_CASE(4, 2)
  DMM_CASE(1, 4) DMM_CASE(2, 4) DMM_CASE(3, 4) DMM_CASE(4, 4)
#undef DMM_CASE
  return DMM_ERR_UNSUPPORTED_SHAPE;
}

}  // namespace

int solver_max_rows() { return kMaxRows; }
int solver_max_cols() { return kMaxCols; }

}  // namespace dmm

using namespace dmm;

static size_t bits_bytes(int B, int max_iter, int proj_iter) {
  return align_up((size_t)B * max_iter * (proj_iter > 0 ? proj_iter : 0) * kThreads * sizeof(uint32_t), 256);
}

extern "C" size_t dmm_relax_saved_bytes(int B, int max_iter, int proj_iter) {
  if (B <= 0 || max_iter <= 0) return 256;
  return bits_bytes(B, max_iter, proj_iter) + align_up((size_t)B * max_iter * sizeof(int), 256) + 256;
}

static void split_saved(void* saved, int B, int max_iter, int proj_iter, uint32_t*& bits, int*& sweeps) {
  bits = (uint32_t*)saved;
  sweeps = (int*)((char*)saved + bits_bytes(B, max_iter, proj_iter));
}

extern "C" int dmm_relax_solve(const float* mat, const float* prop_score, int B, int P, int O, const int* n_prop,
                               const int* n_tmpl, int max_iter, int proj_iter, float lr, int negate, int pad_rule,
                               int is_test, float* R, float* X_final, float* Bmat, float* logic,
                               float* match_score, float* det_score, int* n_list, float* xlist, float* cost,
                               void* saved, void* stream) {
  if (B < 0 || P < 0 || O < 0 || max_iter < 0 || proj_iter < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || O == 0) return DMM_OK;
  if (!mat && P > 0) return DMM_ERR_INVALID_ARGUMENT;
  const int MS = pad_rule ? (P > O + 1 ? P : O + 1) : P;
  if (MS <= 0) return DMM_ERR_INVALID_ARGUMENT;
  SolveParams kp;
  kp.mat = mat; kp.score = prop_score; kp.B = B; kp.P = P; kp.O = O; kp.MS = MS;
  kp.n_prop = n_prop; kp.n_tmpl = n_tmpl; kp.max_iter = max_iter; kp.proj_iter = proj_iter; kp.lr = lr;
  kp.negate = negate; kp.pad_rule = pad_rule; kp.is_test = is_test;
  kp.R = R; kp.Xf = X_final; kp.Bm = Bmat; kp.logic = logic; kp.ms = match_score; kp.ds = det_score;
  kp.n_list = n_list; kp.xlist = xlist; kp.cost = cost;
  kp.saved_bits = nullptr; kp.saved_sweeps = nullptr;
  if (saved) split_saved(saved, B, max_iter, proj_iter, kp.saved_bits, kp.saved_sweeps);
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_shape(O, MS, [&](auto nr, auto cpl) {
    relax_solve_kernel<decltype(nr)::value, decltype(cpl)::value><<<B, kThreads, 0, st>>>(kp);
    return check_launch();
  });
}

extern "C" int dmm_relax_solve_bwd(const float* g_R, const float* g_Xfinal, const float* g_Bmat,
                                   const float* g_match_score, const float* g_det_score, const float* mat,
                                   const float* prop_score, const float* R, const float* logic, const int* n_list,
                                   const void* saved, int B, int P, int O, const int* n_prop, const int* n_tmpl,
                                   int max_iter, int proj_iter, float lr, int negate, int pad_rule, float* g_mat,
                                   float* g_prop_score, void* stream) {
  if (B < 0 || P < 0 || O < 0 || max_iter < 0 || proj_iter < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || O == 0 || P == 0) return DMM_OK;
  if (!mat || !R || !logic || !n_list || !saved || !g_mat) return DMM_ERR_INVALID_ARGUMENT;
  const int MS = pad_rule ? (P > O + 1 ? P : O + 1) : P;
  SolveBwdParams kp;
  kp.gR = g_R; kp.gXf = g_Xfinal; kp.gBm = g_Bmat; kp.gms = g_match_score; kp.gds = g_det_score;
  kp.mat = mat; kp.score = prop_score; kp.R = R; kp.logic = logic; kp.n_list = n_list;
  uint32_t* bits; int* sweeps;
  split_saved(const_cast<void*>(saved), B, max_iter, proj_iter, bits, sweeps);
  kp.saved_bits = bits; kp.saved_sweeps = sweeps;
  kp.B = B; kp.P = P; kp.O = O; kp.MS = MS; kp.n_prop = n_prop; kp.n_tmpl = n_tmpl;
  kp.max_iter = max_iter; kp.proj_iter = proj_iter; kp.lr = lr; kp.negate = negate; kp.pad_rule = pad_rule;
  kp.g_mat = g_mat; kp.g_score = g_prop_score;
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_shape(O, MS, [&](auto nr, auto cpl) {
    relax_solve_bwd_kernel<decltype(nr)::value, decltype(cpl)::value><<<B, kThreads, 0, st>>>(kp);
    return check_launch();
  });
}
