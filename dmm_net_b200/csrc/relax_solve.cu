// K3 -- relaxed matching solver + assignment head, forward and backward.
//
// Reference: dmm/modules/submodules/relax_match.py:36-105 (greedy init :45-55, gradient step :68-71,
// Dykstra sweep :73-87 with project_col :21-34 and project_row :9-19, inner exit :88-89, outer exit :96-98)
// and the [O x m] part of dmm/modules/match_model.py:93-148 (pad rule :109-113, mean of X_list :121,
// logic mask :125-130, match_score :146, det_score :147).
//
// One warp per problem.  Lane l owns columns l, l+32, ... of all rows, so column sums are lane-local
// (sequential over rows, the reference's order) and row sums are xor-butterflies whose result is bit-identical
// in every lane (no divergence on the exit tests).  X, the three Dykstra increments and X_start live in
// registers for the whole solve; C and the running sum of iterates sit in shared memory (touched once per
// outer step).  All arithmetic mirrors the reference's fp32 op order with explicit *_rn intrinsics (no FMA
// contraction); both early exits are evaluated on the device with the reference's exact-equality tests.
#include "common.cuh"

namespace dmm {
namespace {

constexpr int kWarpsPerCta = 4;
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
  unsigned long long* saved_bits;  // [B][max_iter*proj_iter][32]
  int* saved_sweeps;               // [B][max_iter]
};

__device__ __forceinline__ void problem_dims(const int* n_prop, const int* n_tmpl, int b, int P, int O,
                                             int pad_rule, int& n, int& np, int& m) {
  np = n_prop ? clampi(n_prop[b], 0, P) : P;
  n = n_tmpl ? clampi(n_tmpl[b], 0, O) : O;
  m = (pad_rule && np <= n) ? n + 1 : np;
}

template <int NR, int CPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32) relax_solve_kernel(const SolveParams p) {
  constexpr int W = CPL * 32;
  __shared__ float sm[kWarpsPerCta][2][NR * W];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kWarpsPerCta + warp;
  if (b >= p.B) return;  // warps are independent: no block-level barrier below
  float* Cs = sm[warp][0];
  float* Rs = sm[warp][1];

  int n, np, m;
  problem_dims(p.n_prop, p.n_tmpl, b, p.P, p.O, p.pad_rule, n, np, m);
  const float nf = (float)n, mf = (float)m;
  bool cv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) cv[j] = lane + 32 * j < m;

  const long long obase = (long long)b * p.O * p.MS;
  auto write_mat = [&](float* dst, int r, int j, float v) {
    const int col = lane + 32 * j;
    if (dst && col < p.MS) dst[obase + (long long)r * p.MS + col] = v;
  };
  if (n == 0 || m == 0) {  // nothing to match: define every output as zero
    for (int r = 0; r < p.O; ++r) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        write_mat(p.R, r, j, 0.f); write_mat(p.Xf, r, j, 0.f); write_mat(p.Bm, r, j, 0.f); write_mat(p.logic, r, j, 0.f);
      }
      if (lane == 0) {
        if (p.ms) p.ms[(long long)b * p.O + r] = 0.f;
        if (p.ds) p.ds[(long long)b * p.O + r] = 0.f;
      }
    }
    if (lane == 0 && p.n_list) p.n_list[b] = 0;
    return;
  }

  // ---- load the cost -----------------------------------------------------------------------------------
  const float* mat = p.mat + (long long)b * p.O * p.P;
  float fill = -INFINITY;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      float c = 0.f;
      if (r < n && col < np) {
        c = mat[r * p.P + col];
        if (p.negate) c = -c;
      }
      Cs[r * W + col] = c;
      if (r < n && cv[j]) fill = fmaxf(fill, c);
    }
  }
  fill = warp_max(fill);  // C.max() of the padded cost (relax_match.py:51)
  __syncwarp();

  // ---- greedy start (relax_match.py:45-55) -----------------------------------------------------------
  float x[NR][CPL], q0[NR][CPL], q1[NR][CPL], q2[NR][CPL];
  {
    int best[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      float bv = INFINITY;
      best[j] = 0;
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < n) {
          const float c = Cs[r * W + lane + 32 * j];
          if (c < bv) { bv = c; best[j] = r; }  // first minimum wins
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float bv = INFINITY;
      int bi = 0x7fffffff;
      if (r < n) {
#pragma unroll
        for (int j = 0; j < CPL; ++j)
          if (cv[j]) {
            const float kept = best[j] == r ? Cs[r * W + lane + 32 * j] : fill;
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
        x[r][j] = (r < n && lane + 32 * j == bi) ? 1.f : 0.f;
        q0[r][j] = q1[r][j] = q2[r][j] = 0.f;
        Rs[r * W + lane + 32 * j] = x[r][j];  // sum(X_list) starts as 0 + X0 == X0
      }
    }
  }
  int L = 1;
  float* xl = p.xlist ? p.xlist + (long long)b * (p.max_iter + 1) * p.O * p.MS : nullptr;
  float* costv = p.cost ? p.cost + (long long)b * (p.max_iter + 1) : nullptr;
  auto record_iterate = [&](int slot) {
    if (!xl) return;
    float* dst = xl + (long long)slot * p.O * p.MS;
#pragma unroll
    for (int r = 0; r < NR; ++r)
      if (r < p.O) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int col = lane + 32 * j;
          if (col < p.MS) dst[r * p.MS + col] = (r < n && cv[j]) ? x[r][j] : 0.f;
        }
      }
  };
  record_iterate(0);
  if (costv && lane == 0) costv[0] = 0.f;

  unsigned long long* sbits = p.saved_bits ? p.saved_bits + (long long)b * p.max_iter * p.proj_iter * 32 : nullptr;
  int* ssweeps = p.saved_sweeps ? p.saved_sweeps + (long long)b * p.max_iter : nullptr;

  float cost_prev = 0.f;
  for (int it = 0; it < p.max_iter; ++it) {
    // ---- gradient step, cost, record (relax_match.py:69-71) ------------------------------------------
    float c2 = 0.f;
#pragma unroll
    for (int r = 0; r < NR; ++r)
      if (r < n) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int col = lane + 32 * j;
          const float c = Cs[r * W + col];
          x[r][j] = __fsub_rn(x[r][j], __fmul_rn(p.lr, c));
          const float xc = __fmul_rn(x[r][j], c);
          c2 = __fadd_rn(c2, __fmul_rn(xc, xc));
          Rs[r * W + col] = __fadd_rn(Rs[r * W + col], x[r][j]);
        }
      }
    const float cost_cur = __fsqrt_rn(warp_sum(c2));
    record_iterate(L);
    if (costv && lane == 0) costv[L] = cost_cur;
    ++L;

    // ---- Dykstra sweeps (relax_match.py:73-89) ----------------------------------------------------------
    int sweeps = 0;
    for (int js = 0; js < p.proj_iter; ++js) {
      float xs[NR][CPL];
      float cs[CPL];
      unsigned long long sb = 0ull;
#pragma unroll
      for (int j = 0; j < CPL; ++j) cs[j] = 0.f;
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < n) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            xs[r][j] = x[r][j];
            const float a = __fadd_rn(x[r][j], q0[r][j]);
            const float y = fmaxf(a, 0.f);                      // {X >= 0}
            q0[r][j] = __fsub_rn(a, y);
            const float bb = __fadd_rn(y, q1[r][j]);
            x[r][j] = bb;
            cs[j] = __fadd_rn(cs[j], bb);                       // column sum, rows in order
            if (a > 0.f) sb |= 1ull << (r * CPL + j);
          }
        }
      bool keep[CPL];
      float tcol[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        keep[j] = cs[j] <= 1.f;                                 // {col sums <= 1}: only violators move
        tcol[j] = __fdiv_rn(__fsub_rn(cs[j], 1.f), nf);
        if (keep[j]) sb |= 1ull << (NR * CPL + j);
      }
      float rs[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < n) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float bb = x[r][j];
            const float y1 = keep[j] ? bb : __fsub_rn(bb, tcol[j]);
            q1[r][j] = __fsub_rn(bb, y1);
            const float cc = __fadd_rn(y1, q2[r][j]);
            x[r][j] = cc;
            acc = __fadd_rn(acc, cc);
          }
          rs[r] = acc;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (r < n) rs[r] = __fadd_rn(rs[r], __shfl_xor_sync(0xffffffffu, rs[r], o));
      }
      bool changed = false;
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < n) {
          const float u = __fdiv_rn(__fsub_rn(rs[r], 1.f), mf);  // {row sums == 1}
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float cc = x[r][j];
            const float y2 = cv[j] ? __fsub_rn(cc, u) : 0.f;
            q2[r][j] = __fsub_rn(cc, y2);
            x[r][j] = y2;
            const float d = __fsub_rn(y2, xs[r][j]);
            changed |= __fmul_rn(d, d) != 0.f;                    // ||X - X_start|| == 0  <=>  every square is 0
          }
        }
      if (sbits) sbits[((long long)it * p.proj_iter + js) * 32 + lane] = sb;
      ++sweeps;
      if (!__any_sync(0xffffffffu, changed)) break;
    }
    if (ssweeps && lane == 0) ssweeps[it] = sweeps;
    if (cost_prev == cost_cur) break;                             // relax_match.py:96
    cost_prev = cost_cur;
  }
  __syncwarp();

  // ---- head: R = mean(X_list), logic, Bmat, scores (match_model.py:121-147) ----------------------------
  const float Lf = (float)L;
  const float* score = p.score ? p.score + (long long)b * p.P : nullptr;
  for (int r = 0; r < p.O; ++r) {
    float Rv[CPL], top = -INFINITY;
    const bool rv = r < n;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      Rv[j] = (rv && cv[j]) ? __fdiv_rn(Rs[(r < NR ? r : 0) * W + col], Lf) : 0.f;
      if (rv && cv[j]) top = fmaxf(top, Rv[j]);
    }
    top = warp_max(top);
    float best = -INFINITY;
    double det = 0.0;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      float lg = 0.f, bm = 0.f;
      if (rv && cv[j]) {
        lg = p.is_test ? (Rv[j] == top ? 1.f : 0.f) : (Rv[j] > 0.01f ? 1.f : 0.f);
        bm = __fmul_rn(Rv[j], lg);
        const float c = Cs[(r < NR ? r : 0) * W + col];
        const float simv = -c;                                  // (-cost_matrix), match_model.py:146
        best = fmaxf(best, __fmul_rn(fminf(fmaxf(Rv[j], 0.f), 1.f), simv));
        if (score && col < np) det += (double)__fmul_rn(score[col], bm);
      }
      write_mat(p.R, r, j, Rv[j]);
      write_mat(p.logic, r, j, lg);
      write_mat(p.Bm, r, j, bm);
    }
    // X_final needs the register copy: only rows < NR exist
    best = warp_max(best);
    det = warp_sum(det);
    if (lane == 0) {
      if (p.ms) p.ms[(long long)b * p.O + r] = rv ? best : 0.f;
      if (p.ds) p.ds[(long long)b * p.O + r] = rv ? (float)det : 0.f;
    }
  }
  if (p.Xf) {
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (r < p.O) write_mat(p.Xf, r, j, (r < n && cv[j]) ? x[r][j] : 0.f);
  }
  if (lane == 0 && p.n_list) p.n_list[b] = L;
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
struct SolveBwdParams {
  const float *gR, *gXf, *gBm, *gms, *gds;
  const float *mat, *score, *R, *logic;
  const int* n_list;
  const unsigned long long* saved_bits;
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

template <int NR, int CPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32) relax_solve_bwd_kernel(const SolveBwdParams p) {
  constexpr int W = CPL * 32;
  __shared__ float sm[kWarpsPerCta][2][NR * W];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kWarpsPerCta + warp;
  if (b >= p.B) return;
  float* gRl = sm[warp][0];  // d loss / d X_list[k] (identical for every k): gR / L
  float* gC = sm[warp][1];   // accumulated d loss / d C
  int n, np, m;
  problem_dims(p.n_prop, p.n_tmpl, b, p.P, p.O, p.pad_rule, n, np, m);
  float* g_mat = p.g_mat + (long long)b * p.O * p.P;
  float* g_score = p.g_score ? p.g_score + (long long)b * p.P : nullptr;
  const int L = p.n_list[b];
  if (n == 0 || m == 0 || L <= 0) {
    for (int i = lane; i < p.O * p.P; i += 32) g_mat[i] = 0.f;
    if (g_score) for (int i = lane; i < p.P; i += 32) g_score[i] = 0.f;
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
  float gsc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) gsc[j] = 0.f;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (r >= p.O) continue;
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
      float g = 0.f, gdir = 0.f;
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
          gdir = gms * fminf(fmaxf(Rv[j], 0.f), 1.f);
        }
        g = g / Lf;
      }
      gRl[r * W + col] = g;
      gC[r * W + col] = 0.f;
      if (col < p.P) g_mat[r * p.P + col] = (col < np) ? (p.negate ? gdir : -gdir) : 0.f;
    }
  }
  if (g_score) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      if (col < p.P) g_score[col] = col < np ? gsc[j] : 0.f;
    }
  }
  __syncwarp();

  // ---- reverse sweep through the solver ---------------------------------------------------------------
  float gx[NR][CPL], g0[NR][CPL], g1[NR][CPL], g2[NR][CPL];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int col = lane + 32 * j;
      gx[r][j] = (p.gXf && r < n && cv[j]) ? p.gXf[obase + (long long)r * p.MS + col] : 0.f;
      g0[r][j] = g1[r][j] = g2[r][j] = 0.f;
    }
  const unsigned long long* sbits = p.saved_bits + (long long)b * p.max_iter * p.proj_iter * 32;
  const int* ssweeps = p.saved_sweeps + (long long)b * p.max_iter;

  for (int it = L - 2; it >= 0; --it) {
    const int sweeps = ssweeps[it];
    for (int js = sweeps - 1; js >= 0; --js) {
      const unsigned long long sb = sbits[((long long)it * p.proj_iter + js) * 32 + lane];
      // X' = Y2, P2' = Cc - Y2, Y2 = Cc - (rowsum(Cc) - 1)/m
      float rs[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < n) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float gy2 = cv[j] ? gx[r][j] - g2[r][j] : 0.f;
            gx[r][j] = gy2;  // reuse as gY2
            acc += gy2;
          }
          rs[r] = acc;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (r < n) rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], o);
      }
      float hs[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) hs[j] = 0.f;
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < n) {
          const float u = rs[r] / mf;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float gcc = cv[j] ? g2[r][j] + gx[r][j] - u : 0.f;  // d/dCc
            g2[r][j] = gcc;                                              // Cc = Y1 + P2  -> gP2
            const float h = gcc - g1[r][j];                              // P1' = Bb - Y1
            gx[r][j] = h;                                                // reuse as h
            hs[j] += h;
          }
        }
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < n) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const bool keep = (sb >> (NR * CPL + j)) & 1ull;
            const float h = gx[r][j];
            const float gbb = g1[r][j] + (keep ? h : h - hs[j] / nf);    // Y1 = colproj(Bb)
            g1[r][j] = gbb;                                              // Bb = Y0 + P1 -> gP1
            const float t = gbb - g0[r][j];                              // P0' = A - Y0
            const bool pos = (sb >> (r * CPL + j)) & 1ull;
            const float ga = g0[r][j] + (pos ? t : 0.f);                 // Y0 = relu(A)
            g0[r][j] = ga;                                               // A = X + P0 -> gP0
            gx[r][j] = cv[j] ? ga : 0.f;
          }
        }
    }
    // X_g = X_prev - lr*C is both recorded (weight 1/L in R) and fed to the sweeps
#pragma unroll
    for (int r = 0; r < NR; ++r)
      if (r < n) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int col = lane + 32 * j;
          const float g = gx[r][j] + gRl[r * W + col];
          gC[r * W + col] = fmaf(-p.lr, g, gC[r * W + col]);
          gx[r][j] = g;
        }
      }
  }
  // ---- d loss / d mat = direct term (already there) -/+ d loss / d C -----------------------------------
#pragma unroll
  for (int r = 0; r < NR; ++r)
    if (r < n) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int col = lane + 32 * j;
        if (col < np) {
          const float g = gC[r * W + col];
          g_mat[r * p.P + col] += p.negate ? -g : g;
        }
      }
    }
}

template <typename F>
int dispatch_shape(int O, int MS, F&& f) {
  // register tiles: (rows NR, columns per lane CPL); pick the smallest that holds [O x MS]
  if (O <= 4 && MS <= 32) return f(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{});
  if (O <= 4 && MS <= 64) return f(std::integral_constant<int, 4>{}, std::integral_constant<int, 2>{});
  if (O <= 8 && MS <= 64) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 2>{});
  if (O <= 16 && MS <= 64) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 2>{});
  if (O <= 8 && MS <= 128) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 4>{});
  return DMM_ERR_UNSUPPORTED_SHAPE;
}

}  // namespace

int solver_max_rows() { return kMaxRows; }
int solver_max_cols() { return kMaxCols; }

}  // namespace dmm

using namespace dmm;

extern "C" size_t dmm_relax_saved_bytes(int B, int max_iter, int proj_iter) {
  if (B <= 0 || max_iter <= 0) return 256;
  const size_t bits = align_up((size_t)B * max_iter * (proj_iter > 0 ? proj_iter : 0) * 32 * sizeof(unsigned long long), 256);
  const size_t sweeps = align_up((size_t)B * max_iter * sizeof(int), 256);
  return bits + sweeps + 256;
}

static void split_saved(void* saved, int B, int max_iter, int proj_iter, unsigned long long*& bits, int*& sweeps) {
  const size_t nb = align_up((size_t)B * max_iter * (proj_iter > 0 ? proj_iter : 0) * 32 * sizeof(unsigned long long), 256);
  bits = (unsigned long long*)saved;
  sweeps = (int*)((char*)saved + nb);
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
  const int grid = (B + kWarpsPerCta - 1) / kWarpsPerCta;
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_shape(O, MS, [&](auto nr, auto cpl) {
    relax_solve_kernel<decltype(nr)::value, decltype(cpl)::value><<<grid, kWarpsPerCta * 32, 0, st>>>(kp);
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
  unsigned long long* bits; int* sweeps;
  split_saved(const_cast<void*>(saved), B, max_iter, proj_iter, bits, sweeps);
  kp.saved_bits = bits; kp.saved_sweeps = sweeps;
  kp.B = B; kp.P = P; kp.O = O; kp.MS = MS; kp.n_prop = n_prop; kp.n_tmpl = n_tmpl;
  kp.max_iter = max_iter; kp.proj_iter = proj_iter; kp.lr = lr; kp.negate = negate; kp.pad_rule = pad_rule;
  kp.g_mat = g_mat; kp.g_score = g_prop_score;
  const int grid = (B + kWarpsPerCta - 1) / kWarpsPerCta;
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_shape(O, MS, [&](auto nr, auto cpl) {
    relax_solve_bwd_kernel<decltype(nr)::value, decltype(cpl)::value><<<grid, kWarpsPerCta * 32, 0, st>>>(kp);
    return check_launch();
  });
}
