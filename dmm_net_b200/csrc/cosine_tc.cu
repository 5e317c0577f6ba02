// K2-TC -- the feature-similarity contraction on the 5th-generation tensor cores (tcgen05 + TMEM), TMA-fed.
//
// Reference: dmm/utils/match_helper.py:51-64 (F.cosine_similarity between every template and proposal feature) for the
// one-template-set case the reference runs (dmm/modules/dmm_model.py:44).  Same outputs as cosine_fwd_kernel
// (cosine.cu): cos[b,o,p] = <q_o,k_p> / (max(|q_o|,eps) * max(|k_p|,eps)), 0 for padding rows / columns.
//
// Why a tensor-core kernel for 0.5 MFLOP per match: the SIMT kernel is latency-bound (100 ns per match, 1.2 TB/s);
// with the contraction on tcgen05 the kernel is a TMA stream of the (P+O)*D*4 = 122 880 feature bytes per match.
// Parity needs fp32-grade dot products, so the contraction is 3xTF32: every fp32 operand is split on chip into
// hi = x & 0xffffe000 (exactly a TF32 number) and lo = tf32(x - hi); D += lo*hi + hi*lo + hi*hi leaves a relative error
// of ~2^-21 per product (the dropped lo*lo term) with fp32 accumulation in TMEM.
//
// Persistent, warp-specialised, one CTA per SM; a work item is a PAIR of problems sharing one 128 x 32 accumulator
// (rows 0-63 / 64-127 = proposals of the two problems, columns 0-15 / 16-31 = their templates; only the two diagonal
// 64 x 16 blocks are read back).  Per 32-float K chunk (one 128-byte swizzle row):
//   warp 10  (1 lane)  TMA: four cp.async.bulk.tensor.2d boxes (SWIZZLE_128B) land the fp32 chunk of both problems'
//                      proposals [64 x 32] and templates [16 x 32] in a slot of the 8-deep raw ring;
//   warps 0-7          split pass for the proposals (the M operand): thread = one row; LDS.128 (conflict-free thanks to
//                      the swizzle) -> hi / lo go to TENSOR MEMORY with tcgen05.st (row = TMEM lane, K element = column),
//                      the sum of squares of the row stays in a register (the norms ride along for free);
//   warps 8-9          split pass for the templates (the N operand): hi / lo into a small shared-memory operand ring in
//                      the K-major SWIZZLE_128B canonical layout, fence.proxy.async;
//                      (warps 0-3 + 8 take the even chunks, warps 4-7 + 9 the odd ones)
//   warp 11  (1 lane)  12 x tcgen05.mma.kind::tf32 M128 N32 K8, A from TMEM, B from shared memory (4 K steps x
//                      {lo*hi, hi*lo, hi*hi}); tcgen05.commit frees the operand slot / publishes the accumulator;
//   warps 12-15        epilogue: tcgen05.ld 32x32b.x16 of the own problem's 16 columns, divide by the norms, coalesced
//                      stores of cos[b][o][p].  Two accumulator buffers decouple it from the next pair.
// Why A lives in TMEM: with both operands in shared memory the kernel was SHARED-MEMORY-BANDWIDTH bound -- every K8 MMA
// re-reads its whole 128-row A slice, 60 KB of operand fetches per chunk on top of 60 KB of split-pass traffic and the
// 20 KB the TMA writes (measured 0.74 us per chunk; 0.38 us with the MMAs removed).  A in TMEM takes the operand
// fetches of the big operand off the shared-memory port.
// The tensor core's fp32 accumulation truncates when it aligns addends (measured: one accumulator over all 64 K steps
// leaves 1e-5 on cos; the error grows with the number of chained MMAs), so K is cut into 4 groups that accumulate in
// separate TMEM columns and are added with round-to-nearest fp32 in the epilogue (measured 3e-6).
// TMEM map (512 columns): [0,256) two accumulator buffers x 4 K groups x 32 columns; [256,512) four operand slots x
// (32 columns hi + 32 columns lo).
// Rows 50..63 of a box are the next problem's first proposals (2-D tensor map over [B*P][D]); they only ever feed
// accumulator rows nobody reads.  The TMA zero-fills past the end of the tensor and past D.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tcgen05.cuh"

namespace dmm {
namespace {

using namespace tc;

constexpr int kStages = 8;                      // raw fp32 ring: what the TMA keeps in flight (8 x 20 KB per SM)
constexpr int kOpStages = 4;                    // operand ring: A hi/lo in TMEM, B hi/lo in shared memory
constexpr int kKC = 32;                         // floats per K chunk (128-byte swizzle row)
constexpr int kRowsA = 128, kRowsB = 32;        // accumulator M (2 x 64 proposals), N (2 x 16 templates)
constexpr int kPadP = 64, kPadO = 16;
constexpr uint32_t kBytesA = kRowsA * kKC * 4;  // 16 KB
constexpr uint32_t kBytesB = kRowsB * kKC * 4;  // 4 KB
constexpr uint32_t kStageBytes = kBytesA + kBytesB;           // raw slot (both operands): 20 KB
constexpr uint32_t kOpBytes = 2 * kBytesB;                    // B hi + lo: 8 KB
constexpr int kConvWarps = 10;                                // two groups of (4 proposal warps + 1 template warp)
constexpr int kGroupWarps = kConvWarps / 2;
constexpr int kTmaWarp = 10, kMmaWarp = 11, kEpiWarp0 = 12;   // epilogue warps 12-15: warp % 4 = TMEM lane quarter
constexpr int kThreadsTc = 16 * 32;
constexpr size_t kDynSmem = (size_t)kStages * kStageBytes + (size_t)kOpStages * kOpBytes + 1024;   // 193 KB
constexpr int kGroups = 4;                                    // K is accumulated in 4 separate TMEM accumulators ...
constexpr uint32_t kAccCols = kGroups * kRowsB;               // ... of 32 columns each, summed in fp32 by the epilogue
constexpr uint32_t kOpCol0 = 2 * kAccCols;                    // first TMEM column of the A operand ring
constexpr uint32_t kOpCols = 2 * kKC;                         // hi + lo columns per slot
constexpr uint32_t kTmemCols = 512;

#ifdef DMM_K2_TRACE   // debug build only: per-chunk clock64 stamps of CTA 0 (scripts/prof_k2.py --trace)
#define K2_STAMP(slot, g) do { if (blockIdx.x == 0 && (g) < 128 && p.trace && (threadIdx.x & 31) == 0) p.trace[(slot) * 128 + (g)] = clock64(); } while (0)
#else
#define K2_STAMP(slot, g) do { } while (0)
#endif

struct TcParams {
  long long* trace;
  int B, P, O, D, pairs, nK, per_group;
  const int* n_prop;
  const int* n_tmpl;
  float eps;
  float* cos;
};

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 32
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRowsB >> 3) << 17) | ((uint32_t)(kRowsA >> 4) << 24);

__global__ void __launch_bounds__(kThreadsTc, 1) cosine_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap map_k,
                                                                  const __grid_constant__ CUtensorMap map_q) {
  extern __shared__ uint8_t dyn_raw[];
  __shared__ uint64_t raw_full[kStages], raw_empty[kStages], op_full[kOpStages], op_empty[kOpStages];
  __shared__ uint64_t tmem_full[2], tmem_empty[2], norm_full[2];
  __shared__ float s_knorm2[2][2][kRowsA], s_qnorm2[2][2][kRowsB];   // [buffer][converter group]: partial sums of squares
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t smem_base = smem_u32(dyn_raw);
  uint8_t* const smem_gen = dyn_raw;                                 // generic pointer / shared address of the same byte
  const uint32_t stage0 = (smem_base + 1023u) & ~1023u;              // swizzle atoms need 1024-byte alignment
  const uint32_t op0 = stage0 + kStages * kStageBytes;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&raw_full[s]), 1);
      mbar_init(smem_u32(&raw_empty[s]), kGroupWarps);
    }
    for (int l = 0; l < kOpStages; ++l) {
      mbar_init(smem_u32(&op_full[l]), kGroupWarps);
      mbar_init(smem_u32(&op_empty[l]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 128);
      mbar_init(smem_u32(&norm_full[a]), kConvWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kTmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int my_pairs = p.pairs > (int)blockIdx.x ? (p.pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == kTmaWarp) {
    // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    uint32_t g = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int pair = blockIdx.x + it * gridDim.x;
      const int b0 = 2 * pair;
      const bool two = b0 + 1 < p.B;
      for (int kc = 0; kc < p.nK; ++kc, ++g) {
        const uint32_t s = g % kStages;
        if (g >= kStages) mbar_wait(smem_u32(&raw_empty[s]), ((g / kStages) - 1u) & 1u, 101);
        K2_STAMP(0, g);
        const uint32_t bar = smem_u32(&raw_full[s]);
        const uint32_t a_raw = stage0 + s * kStageBytes, b_raw = a_raw + kBytesA;
        if (elect_one()) {
          mbar_expect_tx(bar, (two ? 2u : 1u) * (kBytesA / 2 + kBytesB / 2));
          tma_load_2d(a_raw, &map_k, kc * kKC, b0 * p.P, bar);
          tma_load_2d(b_raw, &map_q, kc * kKC, b0 * p.O, bar);
          if (two) {
            tma_load_2d(a_raw + kBytesA / 2, &map_k, kc * kKC, (b0 + 1) * p.P, bar);
            tma_load_2d(b_raw + kBytesB / 2, &map_q, kc * kKC, (b0 + 1) * p.O, bar);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < kConvWarps) {
    // ===== split pass: fp32 -> (hi, lo) TF32 pair, row norms =====
    // Two groups of (4 proposal warps + 1 template warp) take alternate chunks: a thread's per-chunk latency is mostly
    // synchronisation (two mbarrier waits, tcgen05.wait::st, fences: ~900 cycles measured against ~300 of work for a
    // whole 128-byte row), so two chunks are in flight in the split pass at any time.
    const bool is_a = warp < 8;
    const uint32_t grp = is_a ? (uint32_t)(warp >> 2) : (uint32_t)(warp - 8);
    const int row = is_a ? ((warp & 3) * 32 + lane) : lane;    // operand row; for A also the TMEM lane this warp may write
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    uint32_t g = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int ab = it & 1;
      float nrm = 0.f;
      for (int kc = 0; kc < p.nK; ++kc, ++g) {
        if ((g & 1u) != grp) continue;
        const uint32_t s = g % kStages, l = g % kOpStages;
        mbar_wait(smem_u32(&raw_full[s]), (g / kStages) & 1u, 102);
        if (tid == 0) K2_STAMP(1, g);
        if (tid == 256) K2_STAMP(5, g);
        const uint32_t raw = stage0 + s * kStageBytes + (is_a ? 0u : kBytesA) + row_off;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kOpCol0 + l * kOpCols;
        const uint32_t ob = op0 + l * kOpBytes + row_off;
        uint4 x[8];                                          // the whole 128-byte row: eight independent loads first
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const uint4*>(smem_gen + (raw + ((c ^ sw) << 4) - smem_base));
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
          nrm = fmaf(__uint_as_float(x[c].x), __uint_as_float(x[c].x), nrm);
          nrm = fmaf(__uint_as_float(x[c].y), __uint_as_float(x[c].y), nrm);
          nrm = fmaf(__uint_as_float(x[c].z), __uint_as_float(x[c].z), nrm);
          nrm = fmaf(__uint_as_float(x[c].w), __uint_as_float(x[c].w), nrm);
        }
        if (g >= kOpStages) mbar_wait(smem_u32(&op_empty[l]), ((g / kOpStages) - 1u) & 1u, 103);   // MMAs of chunk g-4 done
        if (tid == 0) K2_STAMP(2, g);
        if (is_a) tc_fence_after();
#pragma unroll
        for (uint32_t h = 0; h < 2; ++h) {                   // 16 K elements at a time keeps the register count down
          uint32_t hi[16], lo[16];
#pragma unroll
          for (uint32_t c4 = 0; c4 < 4; ++c4) {
            const uint4 v = x[4 * h + c4];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hi[4 * c4 + e] = w[e] & 0xffffe000u;
              lo[4 * c4 + e] = __float_as_uint(__fsub_rn(__uint_as_float(w[e]), __uint_as_float(hi[4 * c4 + e]))) & 0xffffe000u;
            }
          }
          if (is_a) {
            tmem_st16(taddr + 16u * h, hi);
            tmem_st16(taddr + kKC + 16u * h, lo);
          } else {
#pragma unroll
            for (uint32_t c4 = 0; c4 < 4; ++c4) {
              const uint32_t off = ((4 * h + c4) ^ sw) << 4;
              *reinterpret_cast<uint4*>(smem_gen + (ob + off - smem_base)) =
                  make_uint4(hi[4 * c4], hi[4 * c4 + 1], hi[4 * c4 + 2], hi[4 * c4 + 3]);
              *reinterpret_cast<uint4*>(smem_gen + (ob + kBytesB + off - smem_base)) =
                  make_uint4(lo[4 * c4], lo[4 * c4 + 1], lo[4 * c4 + 2], lo[4 * c4 + 3]);
            }
          }
        }
        if (is_a) {
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        } else {
          fence_proxy_async();                                // generic-proxy stores -> visible to the tensor core's async proxy
        }
        __syncwarp();
        if (lane == 0) {                                      // one arrival per warp (per-thread arrivals serialise on the barrier)
          mbar_arrive(smem_u32(&raw_empty[s]));               // raw slot consumed (values are in registers / operand ring)
          mbar_arrive(smem_u32(&op_full[l]));
          if (tid == 0) K2_STAMP(3, g);
          if (tid == 256) K2_STAMP(6, g);
        }
      }
      // hand the partial sum of squares to the epilogue (buffer `ab` is free once the epilogue of pair it-2 has arrived)
      if (it >= 2) mbar_wait(smem_u32(&tmem_empty[ab]), ((it >> 1) - 1) & 1, 104);
      if (is_a) s_knorm2[ab][grp][row] = nrm; else s_qnorm2[ab][grp][row] = nrm;
      mbar_arrive(smem_u32(&norm_full[ab]));
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer (whole warp walks the loop, one elected lane issues) =====
    uint32_t g = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int ab = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&tmem_empty[ab]), ((it >> 1) - 1) & 1, 104);
      tc_fence_after();
      int grp = 0, in_grp = 0;                               // K group -> its own 32 accumulator columns
      for (int kc = 0; kc < p.nK; ++kc, ++g) {
        const uint32_t d_tmem = tmem_base + (uint32_t)ab * kAccCols + (uint32_t)grp * kRowsB;
        const bool first = in_grp == 0;
        if (++in_grp == p.per_group) { in_grp = 0; ++grp; }
        const uint32_t l = g % kOpStages;
        mbar_wait(smem_u32(&op_full[l]), (g / kOpStages) & 1u, 105);
        K2_STAMP(4, g);
        tc_fence_after();
        const uint32_t a_hi = tmem_base + kOpCol0 + l * kOpCols, a_lo = a_hi + kKC;
        const uint32_t b_hi = op0 + l * kOpBytes, b_lo = b_hi + kBytesB;
        const uint64_t db_hi = umma_desc(b_hi), db_lo = umma_desc(b_lo);
        if (elect_one()) {
#pragma unroll
          for (uint32_t k = 0; k < kKC / 8; ++k) {           // 8 TF32 per MMA: 8 TMEM columns of A, 32 bytes (2 address units) of B
            tc_mma_tf32_ts(d_tmem, a_lo + 8 * k, db_hi + 2 * k, kIdesc, !(first && k == 0));
            tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, db_lo + 2 * k, kIdesc, 1u);
            tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, db_hi + 2 * k, kIdesc, 1u);
          }
          tc_commit(smem_u32(&op_empty[l]));                  // operand slot reusable once these MMAs have read it
          if (kc == p.nK - 1) tc_commit(smem_u32(&tmem_full[ab]));   // accumulator complete
        }
        __syncwarp();
        K2_STAMP(7, g);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue =====
    const int q4 = warp & 3;                                  // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;
    const int m = row >> 6, pp = row & 63;                    // member of the pair (warp-uniform), proposal index
    for (int it = 0; it < my_pairs; ++it) {
      const int ab = it & 1;
      const int pair = blockIdx.x + it * gridDim.x;
      const int b = 2 * pair + m;
      mbar_wait(smem_u32(&norm_full[ab]), (it >> 1) & 1, 106);
      mbar_wait(smem_u32(&tmem_full[ab]), (it >> 1) & 1, 107);
      tc_fence_after();
      float dot[kPadO];
#pragma unroll
      for (int o = 0; o < kPadO; ++o) dot[o] = 0.f;
      const int ngroups = (p.nK + p.per_group - 1) / p.per_group;
      for (int grp = 0; grp < ngroups; ++grp) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ab * kAccCols + grp * kRowsB + m * kPadO);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int o = 0; o < kPadO; ++o) dot[o] = __fadd_rn(dot[o], __uint_as_float(v[o]));
      }
      const float kn = fmaxf(__fsqrt_rn(__fadd_rn(s_knorm2[ab][0][row], s_knorm2[ab][1][row])), p.eps);
      float qn[kPadO];
#pragma unroll
      for (int o = 0; o < kPadO; ++o)
        qn[o] = fmaxf(__fsqrt_rn(__fadd_rn(s_qnorm2[ab][0][m * kPadO + o], s_qnorm2[ab][1][m * kPadO + o])), p.eps);
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[ab]));                 // accumulator + norm buffers may be overwritten
      if (b < p.B && pp < p.P) {
        const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
        const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
        float* cb = p.cos + (long long)b * p.O * p.P + pp;
#pragma unroll
        for (int o = 0; o < kPadO; ++o)
          if (o < p.O) cb[o * p.P] = (o < nt && pp < np) ? __fdiv_rn(dot[o], __fmul_rn(qn[o], kn)) : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// [rows][D] fp32, box = [box_rows][32 floats], 128-byte swizzle (the UMMA K-major canonical layout)
bool make_feat_map(CUtensorMap* m, const float* base, long long rows, int D, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)D * 4ull};
  const cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Returns DMM_OK when the tensor-core kernel was launched, -1 when the shape / alignment is outside its envelope (the
// caller then runs the SIMT kernel), or a DMM_ERR_* code.
int cosine_tc_try_launch(const float* tmpl_feat, const float* prop_feat, int B, int T, int P, int O, int D,
                         const int* n_prop, const int* n_tmpl, float eps, float* cos, cudaStream_t st) {
  if (T != 1 || P > kPadP || O > kPadO || D % 4 != 0 || D < 4) return -1;
  if (((uintptr_t)tmpl_feat & 15u) || ((uintptr_t)prop_feat & 15u)) return -1;
  if ((long long)B * P > 0x7fffffffLL) return -1;
  CUtensorMap mk, mq;
  if (!make_feat_map(&mk, prop_feat, (long long)B * P, D, kPadP) || !make_feat_map(&mq, tmpl_feat, (long long)B * O, D, kPadO))
    return -1;
  TcParams kp;
  kp.B = B; kp.P = P; kp.O = O; kp.D = D; kp.pairs = (B + 1) / 2; kp.nK = (D + kKC - 1) / kKC;
  kp.per_group = (kp.nK + kGroups - 1) / kGroups;
  kp.n_prop = n_prop; kp.n_tmpl = n_tmpl; kp.eps = eps; kp.cos = cos;
  kp.trace = nullptr;
#ifdef DMM_K2_TRACE
  static long long* trace_buf = nullptr;
  if (getenv("DMM_K2_TRACE_FILE")) {
    if (!trace_buf) cudaMalloc(&trace_buf, 8 * 128 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 8 * 128 * sizeof(long long), st);
    kp.trace = trace_buf;
  }
#endif
  cudaError_t e = cudaFuncSetAttribute(cosine_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem);
  if (e != cudaSuccess) { set_last_cuda_error((int)e); return DMM_ERR_CUDA; }
  const int grid = kp.pairs < kNumSMs ? kp.pairs : kNumSMs;
  cosine_tc_kernel<<<grid, kThreadsTc, kDynSmem, st>>>(kp, mk, mq);
#ifdef DMM_K2_TRACE
  if (kp.trace) {
    static long long host[8 * 128];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, kp.trace, sizeof(host), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("DMM_K2_TRACE_FILE"), "w")) {
      for (int g = 0; g < 128; ++g) {
        for (int sl = 0; sl < 8; ++sl) fprintf(f, "%lld ", host[sl * 128 + g]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
#endif
  return check_launch();
}

}  // namespace dmm
