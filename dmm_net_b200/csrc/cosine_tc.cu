// K2-TC -- the feature-similarity contraction on the 5th-generation tensor cores (tcgen05 + TMEM), TMA-fed.
//
// Reference: dmm/utils/match_helper.py:51-64 (F.cosine_similarity between every template and proposal feature) for the
// one-template-set case the reference runs (dmm/modules/dmm_model.py:44).  Same outputs as cosine_fwd_kernel
// (cosine.cu): cos[b,o,p] = <q_o,k_p> / (max(|q_o|,eps) * max(|k_p|,eps)), 0 for padding rows / columns.
//
// Why a tensor-core kernel for 0.5 MFLOP per match: the SIMT kernel is latency-bound (100 ns per match, 1.2 TB/s);
// with the contraction on tcgen05 the kernel is a pure TMA stream of the (P+O)*D*4 = 122 880 feature bytes per match.
// Parity needs fp32-grade dot products, so the contraction is 3xTF32: every fp32 operand is split on chip into
// hi = x & 0xffffe000 (exactly a TF32 number) and lo = tf32(x - hi); D += lo*hi + hi*lo + hi*hi leaves a relative error
// of ~2^-21 per product (the dropped lo*lo term) with fp32 accumulation in TMEM.
//
// Persistent, warp-specialised, one CTA per SM; a work item is a PAIR of problems sharing one 128 x 32 accumulator
// (rows 0-63 / 64-127 = proposals of the two problems, columns 0-15 / 16-31 = their templates; only the two diagonal
// 64 x 16 blocks are read back).  Per 32-float K chunk (one 128-byte swizzle row):
//   warp 10  (1 lane)  TMA: four cp.async.bulk.tensor.2d boxes (SWIZZLE_128B) land the fp32 chunk of both problems'
//                      proposals [64 x 32] and templates [16 x 32] in a ring slot, already in the K-major canonical layout;
//   warps 0-9          split pass: two threads per operand row (64 bytes each); LDS.128 (conflict-free thanks to the
//                      swizzle) -> hi in place, lo into a slot of the 2-deep lo ring, sum of squares of the half row in a
//                      register (the norms ride along for free); fence.proxy.async + mbarrier arrive;
//   warp 11  (1 lane)  12 x tcgen05.mma.kind::tf32 M128 N32 K8 (4 K steps x {lo*hi, hi*lo, hi*hi}), tcgen05.commit
//                      frees the stage / publishes the accumulator;
//   warps 12-15        epilogue: tcgen05.ld 32x32b.x16 of the own problem's 16 columns, divide by the norms, coalesced
//                      stores of cos[b][o][p].  Two accumulator buffers decouple it from the next pair.
// The tensor core's fp32 accumulation truncates when it aligns addends (measured: one accumulator over all 64 K steps
// leaves 1e-5 on cos; the error grows with the number of chained MMAs), so K is cut into 4 groups that accumulate in
// separate TMEM columns and are added with round-to-nearest fp32 in the epilogue.
// The raw/hi ring is 8 slots deep (160 KB in flight per SM): with 4 slots the kernel ran at the pace of the
// TMA -> split -> MMA -> commit round trip (0.7 us per chunk measured, 3.2 TB/s).
// Rows 50..63 of a box are the next problem's first proposals (2-D tensor map over [B*P][D]); they only ever feed
// accumulator rows nobody reads.  The TMA zero-fills past the end of the tensor and past D.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace dmm {
namespace {

constexpr int kStages = 6;                      // raw fp32 / hi ring: what the TMA keeps in flight (8 x 20 KB per SM)
constexpr int kLoStages = 4;                    // lo ring: lives only from the split pass to the MMAs that read it
constexpr int kKC = 32;                         // floats per K chunk (128-byte swizzle row)
constexpr int kRowsA = 128, kRowsB = 32;        // accumulator M (2 x 64 proposals), N (2 x 16 templates)
constexpr int kPadP = 64, kPadO = 16;
constexpr uint32_t kBytesA = kRowsA * kKC * 4;  // 16 KB
constexpr uint32_t kBytesB = kRowsB * kKC * 4;  // 4 KB
constexpr uint32_t kStageBytes = kBytesA + kBytesB;           // one ring slot (both operands): 20 KB
constexpr int kConvWarps = 10, kConvThreads = kConvWarps * 32;   // two threads per operand row (64 bytes each)
constexpr int kTmaWarp = 10, kMmaWarp = 11, kEpiWarp0 = 12;       // epilogue warps 12-15: warp % 4 = TMEM lane quarter
constexpr int kThreadsTc = 16 * 32;
constexpr size_t kDynSmem = (size_t)(kStages + kLoStages) * kStageBytes + 1024;   // 201 KB
constexpr int kGroups = 4;                                        // K is accumulated in 4 separate TMEM accumulators ...
constexpr uint32_t kAccCols = kGroups * kRowsB;                   // ... of 32 columns each, summed in fp32 by the epilogue
constexpr uint32_t kTmemCols = 2 * kAccCols;                      // double-buffered: 256 columns

struct TcParams {
  int B, P, O, D, pairs, nK, per_group;
  const int* n_prop;
  const int* n_tmpl;
  float eps;
  float* cos;
};

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// K-major, SWIZZLE_128B canonical operand: rows at 128 B, 8-row atoms at 1024 B (SBO), version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024u >> 4) << 32;      // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 32
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRowsB >> 3) << 17) | ((uint32_t)(kRowsA >> 4) << 24);

__global__ void __launch_bounds__(kThreadsTc, 1) cosine_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap map_k,
                                                                  const __grid_constant__ CUtensorMap map_q) {
  extern __shared__ uint8_t dyn_raw[];
  __shared__ uint64_t full_bar[kStages], conv_bar[kStages], empty_bar[kStages], lo_empty[kLoStages];
  __shared__ uint64_t tmem_full[2], tmem_empty[2], norm_full[2];
  __shared__ float s_knorm2[2][2][kRowsA], s_qnorm2[2][2][kRowsB];   // [buffer][row half]: partial sums of squares
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t stage0 = (smem_u32(dyn_raw) + 1023u) & ~1023u;     // swizzle atoms need 1024-byte alignment
  const uint32_t lo0 = stage0 + kStages * kStageBytes;
  uint8_t* const smem_gen = dyn_raw;                                 // generic pointer / shared address of the same byte
  const uint32_t smem_base = smem_u32(dyn_raw);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), kConvWarps);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int l = 0; l < kLoStages; ++l) mbar_init(smem_u32(&lo_empty[l]), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 128);
      mbar_init(smem_u32(&norm_full[a]), kConvThreads);
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
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int it = 0; it < my_pairs; ++it) {
        const int pair = blockIdx.x + it * gridDim.x;
        const int b0 = 2 * pair;
        const bool two = b0 + 1 < p.B;
        for (int kc = 0; kc < p.nK; ++kc, ++g) {
          const uint32_t s = g % kStages;
          if (g >= kStages) mbar_wait(smem_u32(&empty_bar[s]), ((g / kStages) - 1u) & 1u);
          const uint32_t bar = smem_u32(&full_bar[s]);
          const uint32_t a_hi = stage0 + s * kStageBytes, b_hi = a_hi + kBytesA;
          mbar_expect_tx(bar, (two ? 2u : 1u) * (kBytesA / 2 + kBytesB / 2));
          tma_load_2d(a_hi, &map_k, kc * kKC, b0 * p.P, bar);
          tma_load_2d(b_hi, &map_q, kc * kKC, b0 * p.O, bar);
          if (two) {
            tma_load_2d(a_hi + kBytesA / 2, &map_k, kc * kKC, (b0 + 1) * p.P, bar);
            tma_load_2d(b_hi + kBytesB / 2, &map_q, kc * kKC, (b0 + 1) * p.O, bar);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < kConvWarps) {
    // ===== split pass: fp32 -> (hi, lo) TF32 pair, row norms =====
    const bool is_a = warp < 8;
    const int half = is_a ? (tid >> 7) : ((tid - 256) >> 5);   // which 64 bytes of the 128-byte row
    const int row = is_a ? (tid & 127) : lane;                 // operand row owned by this thread
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    uint32_t g = 0;
    for (int it = 0; it < my_pairs; ++it) {
      const int ab = it & 1;
      float nrm = 0.f;
      for (int kc = 0; kc < p.nK; ++kc, ++g) {
        const uint32_t s = g % kStages, l = g % kLoStages;
        mbar_wait(smem_u32(&full_bar[s]), (g / kStages) & 1u);
        if (g >= kLoStages) mbar_wait(smem_u32(&lo_empty[l]), ((g / kLoStages) - 1u) & 1u);
        const uint32_t hi_base = stage0 + s * kStageBytes + (is_a ? 0u : kBytesA) + row_off;
        const uint32_t lo_base = lo0 + l * kStageBytes + (is_a ? 0u : kBytesA) + row_off;
        // all four 16-byte loads first (independent), then the arithmetic, then the eight stores: the thread's chunk
        // latency is the pipeline's period (every converter works on the same chunk), so the loads must overlap
        uint4 x[4];
        uint32_t off[4];
#pragma unroll
        for (uint32_t c4 = 0; c4 < 4; ++c4) {
          off[c4] = (((uint32_t)half * 4u + c4) ^ sw) << 4;
          x[c4] = *reinterpret_cast<const uint4*>(smem_gen + (hi_base + off[c4] - smem_base));
        }
        {
#pragma unroll
          for (uint32_t c4 = 0; c4 < 4; ++c4) {
            const float f0 = __uint_as_float(x[c4].x), f1 = __uint_as_float(x[c4].y), f2 = __uint_as_float(x[c4].z),
                        f3 = __uint_as_float(x[c4].w);
            nrm = fmaf(f0, f0, nrm); nrm = fmaf(f1, f1, nrm); nrm = fmaf(f2, f2, nrm); nrm = fmaf(f3, f3, nrm);
            uint4 h, l;
            h.x = x[c4].x & 0xffffe000u; h.y = x[c4].y & 0xffffe000u; h.z = x[c4].z & 0xffffe000u; h.w = x[c4].w & 0xffffe000u;
            l.x = __float_as_uint(__fsub_rn(f0, __uint_as_float(h.x))) & 0xffffe000u;
            l.y = __float_as_uint(__fsub_rn(f1, __uint_as_float(h.y))) & 0xffffe000u;
            l.z = __float_as_uint(__fsub_rn(f2, __uint_as_float(h.z))) & 0xffffe000u;
            l.w = __float_as_uint(__fsub_rn(f3, __uint_as_float(h.w))) & 0xffffe000u;
            *reinterpret_cast<uint4*>(smem_gen + (hi_base + off[c4] - smem_base)) = h;
            *reinterpret_cast<uint4*>(smem_gen + (lo_base + off[c4] - smem_base)) = l;
          }
        }
        fence_proxy_async();                                  // generic-proxy stores -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&conv_bar[s]));   // one arrival per warp: 320 per-thread arrivals on one
                                                              // mbarrier serialised the whole chunk (measured 1 us / chunk)
      }
      // hand the partial sum of squares to the epilogue (buffer `ab` is free once the epilogue of pair it-2 has arrived)
      if (it >= 2) mbar_wait(smem_u32(&tmem_empty[ab]), ((it >> 1) - 1) & 1);
      if (is_a) s_knorm2[ab][half][row] = nrm; else s_qnorm2[ab][half][row] = nrm;
      mbar_arrive(smem_u32(&norm_full[ab]));
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int it = 0; it < my_pairs; ++it) {
        const int ab = it & 1;
        if (it >= 2) mbar_wait(smem_u32(&tmem_empty[ab]), ((it >> 1) - 1) & 1);
        tc_fence_after();
        for (int kc = 0; kc < p.nK; ++kc, ++g) {
          const int grp = kc / p.per_group;                    // K group -> its own 32 accumulator columns
          const uint32_t d_tmem = tmem_base + (uint32_t)ab * kAccCols + (uint32_t)grp * kRowsB;
          const bool first = kc == grp * p.per_group;
          const uint32_t s = g % kStages, l = g % kLoStages;
          mbar_wait(smem_u32(&conv_bar[s]), (g / kStages) & 1u);
          tc_fence_after();
          const uint32_t a_hi = stage0 + s * kStageBytes, b_hi = a_hi + kBytesA;
          const uint32_t a_lo = lo0 + l * kStageBytes, b_lo = a_lo + kBytesA;
          const uint64_t da_hi = umma_desc(a_hi), da_lo = umma_desc(a_lo), db_hi = umma_desc(b_hi), db_lo = umma_desc(b_lo);
#pragma unroll
          for (uint32_t k = 0; k < kKC / 8; ++k) {           // 8 TF32 per MMA = 32 bytes = 2 descriptor address units
            tc_mma_tf32(d_tmem, da_lo + 2 * k, db_hi + 2 * k, kIdesc, !(first && k == 0));
            tc_mma_tf32(d_tmem, da_hi + 2 * k, db_lo + 2 * k, kIdesc, 1u);
            tc_mma_tf32(d_tmem, da_hi + 2 * k, db_hi + 2 * k, kIdesc, 1u);
          }
          tc_commit(smem_u32(&empty_bar[s]));                 // ring slots reusable once these MMAs have read them
          tc_commit(smem_u32(&lo_empty[l]));
        }
        tc_commit(smem_u32(&tmem_full[ab]));                  // accumulator complete
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue =====
    const int q4 = warp & 3;                                  // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;
    const int m = row >> 6, pp = row & 63;                    // member of the pair (warp-uniform), proposal index
    for (int it = 0; it < my_pairs; ++it) {
      const int ab = it & 1;
      const int pair = blockIdx.x + it * gridDim.x;
      const int b = 2 * pair + m;
      mbar_wait(smem_u32(&norm_full[ab]), (it >> 1) & 1);
      mbar_wait(smem_u32(&tmem_full[ab]), (it >> 1) & 1);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static const EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      ptr = nullptr;
    }
    return (EncodeTiledFn)ptr;
  }();
  return fn;
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
  cudaError_t e = cudaFuncSetAttribute(cosine_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem);
  if (e != cudaSuccess) { set_last_cuda_error((int)e); return DMM_ERR_CUDA; }
  const int grid = kp.pairs < kNumSMs ? kp.pairs : kNumSMs;
  cosine_tc_kernel<<<grid, kThreadsTc, kDynSmem, st>>>(kp, mk, mq);
  return check_launch();
}

}  // namespace dmm
