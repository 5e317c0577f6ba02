// K1 -- pairwise binary-mask IoU (reference: dmm/utils/match_helper.py:9-28 over the [O*P, HW] expansion of
// dmm/modules/match_model.py:83-89; second template set = targets of match_helper.py:30-42).
//
// HBM-bound integer kernel.  Every mask byte is read exactly once:
//   - a CTA owns a pixel slab of ONE problem for ALL of its (<=64) masks;
//   - phase A: each warp streams 512-byte row pieces (one LDG.128 per lane, next chunk's loads already in
//     flight), thresholds >0.5 and turns 128 pixels into 4 bit-plane words with __ballot_sync;
//   - phase B: warp w owns word w of the chunk; lanes are proposals, the O template words are smem
//     broadcasts: popc(a & b) into O x 2 register counters per lane;
//   - per-slab int32 partials go to a workspace (no atomics, no memset), a second tiny kernel adds the
//     slabs and forms  inter / (float(|A|+|B|-inter) + 1e-6f)  -- integer-exact, so bit-equal to the reference.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace dmm {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunkPx = 256;                 // pixels per pipeline step (2 groups of 128)
constexpr int kWords = kChunkPx / 32;         // 8 bit-plane words per row per chunk == kWarps
constexpr int kMaxRows = 64;                  // masks per tile (proposals + templates)
constexpr int kMaxUnits = kMaxRows * 2 / kWarps;  // 512-byte row pieces per warp per chunk (16)
constexpr int kTileO = 16;                    // templates per tile (register counters)
static_assert(kWords == kWarps, "one bit-plane word per warp in phase B");

struct IouParams {
  const float* prop;
  const float* const* prop_ptrs;   // optional per-problem base pointers (ragged per-video tensors); overrides prop/prop_bs
  const float* tmpl;
  const float* tmpl2;
  long long prop_bs, tmpl_bs, tmpl2_bs;
  const int* n_prop;
  const int* n_tmpl;
  int P, O, Otot, HW;
  int PT, OT, n_ptiles, n_otiles;  // tile sizes / counts
  int S, chunks_per_slab, n_chunks;
  int B_items;                     // problems in this launch (the persistent TMA kernel walks S * B_items items)
  int* item_counter;               // zeroed per launch: dynamic work distribution of the persistent TMA kernel
  int* ws;                         // [B][S][cnt]
  int cnt;                         // Otot*P + Otot + P
};

// Unconditional 4-pixel read: the index is clamped into the row so the load never branches.  (Round-1 profile: with
// branched loads ptxas put every LDG behind a scoreboard wait of the previous one -- 16 serialised loads per chunk,
// 26 % of the HBM roofline; back-to-back loads gave 56-70 %.)  Pixels past the end only exist in a row's last
// chunk, which takes the masked path below.
template <bool VEC>
__device__ __forceinline__ float4 load_px4(const float* row, int px, int HW) {
  float4 v;
  if (VEC) {
    v = ld_stream_f4(row + min(px, HW - 4));
  } else {
    const int last = HW - 1;
    v.x = ld_stream_f1(row + min(px + 0, last));
    v.y = ld_stream_f1(row + min(px + 1, last));
    v.z = ld_stream_f1(row + min(px + 2, last));
    v.w = ld_stream_f1(row + min(px + 3, last));
  }
  return v;
}

template <bool VEC, int TO>   // TO: template-row counters per lane (ocnt <= TO <= kTileO), loops run unpredicated
__global__ void __launch_bounds__(kThreads, 2) mask_iou_partial_kernel(const IouParams p) {
  // Padding rows (beyond this problem's P+O, or beyond n_prop/n_tmpl) alias a real row: their bits are garbage but
  // the finalize kernel never reads the counters of padding rows, so the hot loop carries no per-row predicate.
  __shared__ const float* row_ptr[kMaxRows];
  __shared__ uint32_t bits[2][2][kMaxRows + kTileO][4];   // [buffer][128-px group][row][word]: one STS.128 per row piece
  __shared__ int red[kTileO * kMaxRows + kMaxRows];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ptile = blockIdx.z % p.n_ptiles, otile = blockIdx.z / p.n_ptiles;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  const int p0 = ptile * p.PT, o0 = otile * p.OT;
  const int pcnt = min(p.PT, p.P - p0), ocnt = min(p.OT, p.Otot - o0);
  const int rows = pcnt + ocnt;
  const float* prop_b = p.prop_ptrs ? p.prop_ptrs[b] : p.prop + (long long)b * p.prop_bs;

  if (tid < kMaxRows) {
    const float* ptr = nullptr;
    if (tid < pcnt) {
      if (p0 + tid < np) ptr = prop_b + (long long)(p0 + tid) * p.HW;
    } else if (tid < rows) {
      const int t = o0 + tid - pcnt;
      if (t < p.O) {
        if (t < nt) ptr = p.tmpl + (long long)b * p.tmpl_bs + (long long)t * p.HW;
      } else if (t - p.O < nt) {
        ptr = p.tmpl2 + (long long)b * p.tmpl2_bs + (long long)(t - p.O) * p.HW;
      }
    }
    row_ptr[tid] = ptr ? ptr : p.tmpl + (long long)b * p.tmpl_bs;  // any readable row of this problem
  }
  for (int i = tid; i < kTileO * kMaxRows + kMaxRows; i += kThreads) red[i] = 0;
  __syncthreads();

  const int c0 = s * p.chunks_per_slab;
  const int c1 = min(c0 + p.chunks_per_slab, p.n_chunks);
  const int tail_chunk = (p.HW % kChunkPx) ? p.n_chunks - 1 : -1;   // the only chunk with pixels past the row end

  int acc[TO][2];
#pragma unroll
  for (int o = 0; o < TO; ++o) acc[o][0] = acc[o][1] = 0;
  int area0 = 0, area1 = 0;  // popcount of row `lane` and row `lane+32` (covers proposals AND templates)

  // this warp's 16 row pieces per chunk: unit u = warp + 8k -> (row u>>1, 128-pixel group u&1); 8*16 = all 64 rows
  const int lane_px = lane * 4;
  const int grp_b = warp >> 2, word_b = warp & 3;   // phase B: this warp's bit-plane word of the chunk
  float4 v[kMaxUnits];

#define DMM_ISSUE(chunk)                                                                        \
  {                                                                                             \
    const int base_ = (chunk) * kChunkPx + lane_px;                                             \
    _Pragma("unroll") for (int k = 0; k < kMaxUnits; ++k) {                                     \
      const int u = warp + kWarps * k;                                                          \
      v[k] = load_px4<VEC>(row_ptr[u >> 1], base_ + (u & 1) * 128, p.HW);                       \
    }                                                                                           \
  }

  if (c0 < c1) DMM_ISSUE(c0);
  for (int c = c0; c < c1; ++c) {
    const int buf = (c - c0) & 1;
    // ---- phase A: threshold + ballot -> bit planes --------------------------------------------------------
    // bit i of word j is pixel 4*i+j of the 128-pixel group: a fixed permutation shared by all rows, and
    // AND/popcount do not care about bit order.
    if (c != tail_chunk) {
#pragma unroll
      for (int k = 0; k < kMaxUnits; ++k) {
        const int u = warp + kWarps * k;
        uint4 w;
        w.x = __ballot_sync(0xffffffffu, v[k].x > 0.5f);
        w.y = __ballot_sync(0xffffffffu, v[k].y > 0.5f);
        w.z = __ballot_sync(0xffffffffu, v[k].z > 0.5f);
        w.w = __ballot_sync(0xffffffffu, v[k].w > 0.5f);
        // every lane holds the same four words: an unconditional same-address STS.128 is one wavefront, no branch
        *reinterpret_cast<uint4*>(&bits[buf][u & 1][u >> 1][0]) = w;
      }
    } else {
      const int base = c * kChunkPx + lane_px;
#pragma unroll
      for (int k = 0; k < kMaxUnits; ++k) {
        const int u = warp + kWarps * k;
        const int px = base + (u & 1) * 128;
        uint4 w;
        w.x = __ballot_sync(0xffffffffu, px + 0 < p.HW && v[k].x > 0.5f);
        w.y = __ballot_sync(0xffffffffu, px + 1 < p.HW && v[k].y > 0.5f);
        w.z = __ballot_sync(0xffffffffu, px + 2 < p.HW && v[k].z > 0.5f);
        w.w = __ballot_sync(0xffffffffu, px + 3 < p.HW && v[k].w > 0.5f);
        *reinterpret_cast<uint4*>(&bits[buf][u & 1][u >> 1][0]) = w;
      }
    }
    if (c + 1 < c1) DMM_ISSUE(c + 1);  // next chunk's 16 loads per lane fly during the barrier and phase B
    __syncthreads();
    // ---- phase B: warp w owns word w of the chunk; lanes are proposals (lane, lane+32) ---------------------
    {
      const uint32_t b0 = bits[buf][grp_b][lane][word_b], b1 = bits[buf][grp_b][lane + 32][word_b];
      area0 += __popc(b0);
      area1 += __popc(b1);
#pragma unroll
      for (int o = 0; o < TO; ++o) {   // rows past ocnt hold stale bits: counted, never written out
        const uint32_t a = bits[buf][grp_b][pcnt + o][word_b];   // smem broadcast
        acc[o][0] += __popc(a & b0);
        acc[o][1] += __popc(a & b1);
      }
    }
    // the double-buffered bit planes make a second barrier unnecessary: buffer `buf` is rewritten in
    // iteration c+2, which every warp enters only after the barrier of iteration c+1.
  }
#undef DMM_ISSUE

  // ---- cross-warp reduction of this slab ---------------------------------------------------------------
#pragma unroll
  for (int o = 0; o < TO; ++o) {
    if (o < ocnt) {
      atomicAdd(&red[o * kMaxRows + lane], acc[o][0]);
      atomicAdd(&red[o * kMaxRows + lane + 32], acc[o][1]);
    }
  }
  atomicAdd(&red[kTileO * kMaxRows + lane], area0);
  atomicAdd(&red[kTileO * kMaxRows + lane + 32], area1);
  __syncthreads();

  int* out = p.ws + ((long long)b * p.S + s) * p.cnt;
  for (int i = tid; i < ocnt * pcnt; i += kThreads) {
    const int o = i / pcnt, q = i - o * pcnt;
    out[(o0 + o) * p.P + p0 + q] = red[o * kMaxRows + q];
  }
  int* area_t = out + p.Otot * p.P;
  int* area_p = area_t + p.Otot;
  if (ptile == 0)
    for (int i = tid; i < ocnt; i += kThreads) area_t[o0 + i] = red[kTileO * kMaxRows + pcnt + i];
  if (otile == 0)
    for (int i = tid; i < pcnt; i += kThreads) area_p[p0 + i] = red[kTileO * kMaxRows + i];
}

// =========================================================================================================
// K1, TMA variant: the same computation with the mask tiles staged by the TMA engine.
//   - three tensor maps (proposals, templates, optional targets) describe the [B][rows][HW] fp32 tensors; ONE
//     cp.async.bulk.tensor.3d (SASS UTMALDG) per tensor and chunk copies a [rows x 256 px] box into a 64 KB stage
//     and completes on an mbarrier (expect_tx); pixels past the row end are zero-filled by the TMA (no tail path);
//   - warp-specialised: one producer warp (a single elected lane) runs ahead through a kStages-deep ring guarded
//     by full/empty mbarriers; 16 consumer warps read the landed stage with conflict-free LDS.128, ballot the
//     threshold into bit planes and popcount -- no register staging, no scoreboard coupling to the loads;
//   - consumers synchronise among themselves with a named barrier; the producer never joins it.
// Requires 16-byte aligned rows (HW % 4 == 0) and a single tile (P + Otot <= 64, Otot <= 16): the common case.
// (First attempt, kept in profiles/r1_k1_ncu_tma_v0.txt: per-row cp.async.bulk issued by 32 lanes of a consumer
//  warp -- UBLKCP is a uniform-datapath instruction, the 64 copies per stage serialised on the critical path: 3.5 TB/s.)
// =========================================================================================================
constexpr int kTmaConsWarps = 16;
constexpr int kTmaConsThreads = kTmaConsWarps * 32;
constexpr int kTmaThreads = kTmaConsThreads + 32;                 // + one producer warp
constexpr int kStages = 3;
constexpr int kStageFloats = kMaxRows * kChunkPx;                 // 64 KB per stage
constexpr size_t kTmaDynSmem = (size_t)kStages * kStageFloats * sizeof(float) + 128;
constexpr int kTmaUnits = kMaxRows * 2 / kTmaConsWarps;           // 8 row pieces per warp per chunk

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsThreads) : "memory"); }

// ROWS / STAGES: 64 rows x 3 stages (64 KB each) is the inference shape (P + O <= 64, O <= 16).  Training adds the targets as
// a second template set -- 50 + 10 + 10 = 70 rows, 20 template rows at the headline shape -- which used to fall back to two
// passes that read the proposals twice (1.71x the algorithmic traffic); the wide instantiation (96 rows x 2 stages of
// 96 KB, up to 32 template rows) takes it in ONE pass.
template <int TO, int ROWS = kMaxRows, int STAGES = 3>
__global__ void __launch_bounds__(kTmaThreads, 1)
mask_iou_partial_tma_kernel(const IouParams p, const __grid_constant__ CUtensorMap tm_prop,
                            const __grid_constant__ CUtensorMap tm_tmpl, const __grid_constant__ CUtensorMap tm_tmpl2) {
  constexpr int kStages = STAGES;                                 // shadows the file-level default inside this kernel
  constexpr int kStageFloats = ROWS * kChunkPx;
  constexpr int kTmaUnits = ROWS * 2 / kTmaConsWarps;
  constexpr int kTOcap = TO <= kTileO ? kTileO : 32;              // template rows the reduction buffer holds
  constexpr bool kWide = ROWS > kMaxRows;
  // PERSISTENT: one CTA per SM claims work items (problem b, pixel slab s) from an atomic counter.  The stage ring
  // and its mbarrier phases run on across items, so the producer is already filling the next item's stages while the
  // consumers reduce and write the previous item's counters: no pipeline fill/drain per slab (with one CTA per SM
  // there is no second CTA to hide it).
  constexpr int TH = TO / 2;                                      // template counters per warp half
  extern __shared__ unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint32_t bits[2][2][ROWS + kTileO][4];
  __shared__ int red[kTOcap * kMaxRows + ROWS];
  __shared__ __align__(8) unsigned long long full_bar[kStages];
  __shared__ __align__(8) unsigned long long empty_bar[kStages];
  __shared__ int stage_item[kStages];   // work item of the chunk in each stage (-1: stop), written by the producer
  __shared__ int stage_last[kStages];   // 1 when that chunk is the last of its item

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pcnt = p.P, ocnt = p.Otot;                            // single tile (checked on the host)
  const bool two = p.tmpl2 != nullptr;
  const int n_items = p.S * p.B_items;

  for (int i = tid; i < kTOcap * kMaxRows + ROWS; i += kTmaThreads) red[i] = 0;
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) {
      mbar_init(smem_u32(&full_bar[st]), 1);
      mbar_init(smem_u32(&empty_bar[st]), kTmaConsWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();   // the only CTA-wide barrier: the producer warp never joins another one

  if (warp == kTmaConsWarps) {
    // ---- producer: one elected lane keeps the ring full, across items ------------------------------------
    // Items are claimed dynamically (atomic counter): SMs on the far die stream ~10 % slower, a static split would
    // run at the pace of the slowest SM (measured: 6.9 TB/s static vs 7.2 TB/s with the hardware CTA scheduler).
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(pcnt + ocnt) * kChunkPx * 4u;   // full boxes: OOB pixels are zero-filled
      unsigned g = 0;                                                   // chunks issued by this CTA so far
      while (true) {
        const int item = atomicAdd(p.item_counter, 1);
        if (item >= n_items) {                                          // sentinel stage: tells the consumers to stop
          const unsigned st = g % kStages;
          if (g >= kStages) mbar_wait(smem_u32(&empty_bar[st]), ((g / kStages) - 1u) & 1u);
          stage_item[st] = -1;
          mbar_arrive(smem_u32(&full_bar[st]));
          break;
        }
        const int b = item / p.S, s = item - b * p.S;
        const int c0 = s * p.chunks_per_slab, c1 = min(c0 + p.chunks_per_slab, p.n_chunks);
        for (int c = c0; c < c1; ++c, ++g) {
          const unsigned st = g % kStages;
          if (g >= kStages) mbar_wait(smem_u32(&empty_bar[st]), ((g / kStages) - 1u) & 1u);
          const uint32_t bar = smem_u32(&full_bar[st]);
          const uint32_t dst = smem_u32(stage + (size_t)st * kStageFloats);
          const int px0 = c * kChunkPx;
          stage_item[st] = item;                                        // published by the arrive below (release)
          stage_last[st] = (c == c1 - 1);
          mbar_expect_tx(bar, bytes);
          tma_load_3d(dst, &tm_prop, px0, 0, b, bar);
          tma_load_3d(dst + (uint32_t)pcnt * kChunkPx * 4u, &tm_tmpl, px0, 0, b, bar);
          if (two) tma_load_3d(dst + (uint32_t)(pcnt + p.O) * kChunkPx * 4u, &tm_tmpl2, px0, 0, b, bar);
        }
      }
    }
    return;
  }

  // ---- consumers ---------------------------------------------------------------------------------------------
  const int word = warp & 7, half = warp >> 3;
  const int grp_b = word >> 2, word_b = word & 3;
  const int lane_px = lane * 4;
  unsigned g = 0;                                                 // chunks consumed by this CTA so far
  int acc[TH][2];
#pragma unroll
  for (int o = 0; o < TH; ++o) acc[o][0] = acc[o][1] = 0;
  int area0 = 0, area1 = 0, area2 = 0;
  while (true) {
    const unsigned st = g % kStages, buf = g & 1u;
    mbar_wait(smem_u32(&full_bar[st]), (g / kStages) & 1u);
    const int item = stage_item[st];
    if (item < 0) break;                                          // sentinel: no more work
    const bool last = stage_last[st] != 0;
    const float* sbase = stage + (size_t)st * kStageFloats;
    // ---- phase A: landed stage -> bit planes (8 row pieces per warp) --------------------------------------
    // rows >= pcnt+ocnt of the stage are never written by the TMA: stale bits, counted into counters nobody reads
#pragma unroll
    for (int k = 0; k < kTmaUnits; ++k) {
      const int u = warp + kTmaConsWarps * k;
      const float4 v = *reinterpret_cast<const float4*>(sbase + (u >> 1) * kChunkPx + (u & 1) * 128 + lane_px);
      uint4 w;
      w.x = __ballot_sync(0xffffffffu, v.x > 0.5f);
      w.y = __ballot_sync(0xffffffffu, v.y > 0.5f);
      w.z = __ballot_sync(0xffffffffu, v.z > 0.5f);
      w.w = __ballot_sync(0xffffffffu, v.w > 0.5f);
      *reinterpret_cast<uint4*>(&bits[buf][u & 1][u >> 1][0]) = w;
    }
    // this warp is done with the stage (its loads fed the ballots above): hand it back to the producer.  The
    // __syncwarp orders every lane's reads of the stage and of stage_item/stage_last before lane 0's release.
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&empty_bar[st]));
    consumer_barrier();
    // ---- phase B: 16 warps = 8 words x 2 halves of the template rows ---------------------------------------
    {
      const uint32_t b0 = bits[buf][grp_b][lane][word_b], b1 = bits[buf][grp_b][lane + 32][word_b];
      if (half == 0) {
        area0 += __popc(b0);
        area1 += __popc(b1);
        if (kWide) area2 += __popc(bits[buf][grp_b][min(lane + 64, ROWS + kTileO - 1)][word_b]);   // rows 64..95 (template rows)
      }
#pragma unroll
      for (int o = 0; o < TH; ++o) {
        const uint32_t a = bits[buf][grp_b][pcnt + half * TH + o][word_b];
        acc[o][0] += __popc(a & b0);
        acc[o][1] += __popc(a & b1);
      }
    }
    ++g;
    // double-buffered bit planes: buffer `buf` is rewritten two chunks later, after the next chunk's barrier
    if (!last) continue;

    // ---- item epilogue (the producer keeps loading the next item meanwhile) ----------------------------------
#pragma unroll
    for (int o = 0; o < TH; ++o) {
      const int oo = half * TH + o;
      if (oo < ocnt) {
        atomicAdd(&red[oo * kMaxRows + lane], acc[o][0]);
        atomicAdd(&red[oo * kMaxRows + lane + 32], acc[o][1]);
      }
      acc[o][0] = acc[o][1] = 0;
    }
    if (half == 0) {
      atomicAdd(&red[kTOcap * kMaxRows + lane], area0);
      atomicAdd(&red[kTOcap * kMaxRows + lane + 32], area1);
      if (kWide && lane + 64 < ROWS) atomicAdd(&red[kTOcap * kMaxRows + lane + 64], area2);
    }
    area0 = area1 = area2 = 0;
    consumer_barrier();
    const int b = item / p.S, s = item - b * p.S;
    int* out = p.ws + ((long long)b * p.S + s) * p.cnt;
    for (int i = tid; i < ocnt * pcnt; i += kTmaConsThreads) {
      const int o = i / pcnt, q = i - o * pcnt;
      out[o * p.P + q] = red[o * kMaxRows + q];
    }
    int* area_t = out + p.Otot * p.P;
    int* area_p = area_t + p.Otot;
    for (int i = tid; i < ocnt; i += kTmaConsThreads) area_t[i] = red[kTOcap * kMaxRows + pcnt + i];
    for (int i = tid; i < pcnt; i += kTmaConsThreads) area_p[i] = red[kTOcap * kMaxRows + i];
    consumer_barrier();
    for (int i = tid; i < kTOcap * kMaxRows + ROWS; i += kTmaConsThreads) red[i] = 0;
    // the next epilogue's atomics come after at least one more chunk barrier: the zeroing above is ordered before them
  }
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) -------------------
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

// [B][rows][HW] fp32 with batch stride `bs` elements; box = [1][rows][256 px]
bool make_map(CUtensorMap* m, const float* base, long long bs, int B, int rows, int HW) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)rows, (cuuint64_t)B};
  const cuuint64_t gstr[2] = {(cuuint64_t)HW * 4ull, (cuuint64_t)(B > 1 ? bs : (long long)rows * HW) * 4ull};
  const cuuint32_t box[3] = {(cuuint32_t)kChunkPx, (cuuint32_t)rows, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int TO, int ROWS = kMaxRows, int STAGES = 3>
int launch_tma(const IouParams& kp, const CUtensorMap& mp, const CUtensorMap& mt, const CUtensorMap& mt2, dim3 grid,
               cudaStream_t st) {
  constexpr size_t dyn = (size_t)STAGES * ROWS * kChunkPx * sizeof(float) + 128;
  cudaError_t e = cudaFuncSetAttribute(mask_iou_partial_tma_kernel<TO, ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)dyn);
  if (e != cudaSuccess) { set_last_cuda_error((int)e); return DMM_ERR_CUDA; }
  const long long items = (long long)grid.x * grid.y;   // S * B work items, one persistent CTA per SM walks them
  const int nctas = (int)(items < kNumSMs ? items : kNumSMs);
  mask_iou_partial_tma_kernel<TO, ROWS, STAGES><<<nctas, kTmaThreads, dyn, st>>>(kp, mp, mt, mt2);
  return check_launch();
}

constexpr int kWideRows = 96, kWideTO = 32;     // the wide single-pass tile of the TMA kernel (training: targets ride along)
inline bool wide_shape(int P, int Otot) {
  return P <= kMaxRows && Otot <= kWideTO && P + Otot <= kWideRows && (P + Otot > kMaxRows || Otot > kTileO);
}

struct FinParams {
  const int* ws;
  int S, cnt, B, P, O, Otot;
  const int* n_prop;
  const int* n_tmpl;
  float* iou;
  float* iou2;
  const float* cos;
  float w_cos, w_iou;
  float* sim;
  int* counts;
};

__global__ void __launch_bounds__(256) mask_iou_finalize_kernel(const FinParams p) {
  const long long total = (long long)p.B * p.Otot * p.P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % p.P);
    const int t = (int)((i / p.P) % p.Otot);
    const int b = (int)(i / ((long long)p.P * p.Otot));
    const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
    const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
    const int o = t < p.O ? t : t - p.O;
    float r = 0.f;
    int inter = 0, at = 0, ap = 0;
    if (q < np && o < nt) {
      const int* w = p.ws + (long long)b * p.S * p.cnt;
      for (int s = 0; s < p.S; ++s, w += p.cnt) {
        inter += w[t * p.P + q];
        at += w[p.Otot * p.P + t];
        ap += w[p.Otot * p.P + p.Otot + q];
      }
      // match_helper.py:21-27: union.sum()+1e-6 in fp32, IEEE divide.  The counts are exact integers < 2^24.
      const float uni = __fadd_rn((float)(at + ap - inter), 1e-6f);
      r = __fdiv_rn((float)inter, uni);
    }
    const long long oi = ((long long)b * p.O + o) * p.P + q;
    if (t < p.O) {
      if (p.iou) p.iou[oi] = r;
      if (p.sim) p.sim[oi] = __fadd_rn(__fmul_rn(p.cos[oi], p.w_cos), __fmul_rn(r, p.w_iou));  // match_model.py:90
      if (p.counts) {
        int* c = p.counts + (long long)b * (p.O * p.P + p.O + p.P);
        c[o * p.P + q] = inter;
        if (q == 0) c[p.O * p.P + o] = at;
        if (o == 0) c[p.O * p.P + p.O + q] = ap;
      }
    } else if (p.iou2) {
      p.iou2[oi] = r;
    }
  }
}

struct Plan {
  int PT, OT, n_ptiles, n_otiles, S, n_chunks, chunks_per_slab, cnt, Otot;
};

Plan make_plan(int B, int P, int O, int HW, int two, bool wide = false) {
  Plan pl;
  pl.Otot = O * (two ? 2 : 1);
  pl.OT = pl.Otot < kTileO ? pl.Otot : kTileO;
  pl.PT = P < kMaxRows - pl.OT ? P : kMaxRows - pl.OT;
  if (wide) { pl.OT = pl.Otot; pl.PT = P; }     // one tile of <= 96 rows (TMA kernel only)
  if (pl.OT < 1) pl.OT = 1;
  if (pl.PT < 1) pl.PT = 1;
  pl.n_ptiles = (P + pl.PT - 1) / pl.PT;
  pl.n_otiles = (pl.Otot + pl.OT - 1) / pl.OT;
  pl.n_chunks = (HW + kChunkPx - 1) / kChunkPx;
  if (pl.n_chunks < 1) pl.n_chunks = 1;
  // ~24 work items per SM (the persistent TMA kernel strides over them: imbalance <= one small item; the LDG
  // kernel gets 12 waves of 2 CTAs/SM), but never slabs shorter than 4 chunks
  const long long tiles = (long long)B * pl.n_ptiles * pl.n_otiles;
  long long want = (24LL * kNumSMs + tiles - 1) / tiles;
  long long max_s = pl.n_chunks / 4 > 0 ? pl.n_chunks / 4 : 1;
  long long S = want < 1 ? 1 : (want > max_s ? max_s : want);
  if (S > 65535) S = 65535;
  pl.chunks_per_slab = (int)((pl.n_chunks + S - 1) / S);
  pl.S = (pl.n_chunks + pl.chunks_per_slab - 1) / pl.chunks_per_slab;
  pl.cnt = pl.Otot * P + pl.Otot + P;
  return pl;
}

}  // namespace
}  // namespace dmm

using namespace dmm;

extern "C" size_t dmm_mask_iou_workspace_bytes(int B, int P, int O, int HW, int two_template_sets) {
  if (B <= 0 || P <= 0 || O <= 0 || HW <= 0) return 0;
  const Plan pl = make_plan(B, P, O, HW, two_template_sets);
  size_t bytes = (size_t)B * pl.S * pl.cnt * sizeof(int);
  if (wide_shape(P, pl.Otot)) {                 // which of the two plans runs is decided at launch (alignment, tensor maps)
    const Plan pw = make_plan(B, P, O, HW, two_template_sets, true);
    const size_t bw = (size_t)B * pw.S * pw.cnt * sizeof(int);
    if (bw > bytes) bytes = bw;
  }
  return align_up(bytes, 256) + 256;            // + the persistent kernel's work counter
}

static int run_pairwise(const float* prop, const float* const* prop_ptrs, int ptrs_aligned16, long long prop_bstride,
                        const float* tmpl, long long tmpl_bstride, const float* tmpl2, long long tmpl2_bstride, int B,
                        int P, int O, int HW, const int* n_prop, const int* n_tmpl, float* iou, float* iou2,
                        const float* cos, float w_cos, float w_iou, float* sim, int* counts, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (B < 0 || P < 0 || O < 0 || HW < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || P == 0 || O == 0) return DMM_OK;  // nothing to write
  if (!workspace || (HW > 0 && ((!prop && !prop_ptrs) || !tmpl))) return DMM_ERR_INVALID_ARGUMENT;
  if (sim && !cos) return DMM_ERR_INVALID_ARGUMENT;
  if (iou2 && !tmpl2) return DMM_ERR_INVALID_ARGUMENT;
  if (B > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;
  const int two = tmpl2 != nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  auto aligned16 = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
  const bool vec = (HW % 4 == 0) && (prop_ptrs ? ptrs_aligned16 != 0 : aligned16(prop)) && aligned16(tmpl) && (!two || aligned16(tmpl2)) &&
                   (prop_bstride % 4 == 0) && (tmpl_bstride % 4 == 0) && (!two || tmpl2_bstride % 4 == 0);
  // Staging path: the TMA ring is the default whenever it applies (aligned rows, single tile, tensor maps available);
  // measured on par with / slightly ahead of the LDG pipeline (profiles/README.md).  DMM_K1_IMPL=ldg forces the
  // LDG kernel (read-only environment lookup, used by the A/B parity test).
  const char* impl = getenv("DMM_K1_IMPL");
  const bool tma_ok = vec && !prop_ptrs && !(impl && impl[0] == 'l') && HW >= kChunkPx;
  CUtensorMap mp, mt, mt2;
  bool maps = false;
  if (tma_ok) {
    maps = make_map(&mp, prop, prop_bstride, B, P, HW) && make_map(&mt, tmpl, tmpl_bstride, B, O, HW);
    if (maps && two) maps = make_map(&mt2, tmpl2, tmpl2_bstride, B, O, HW);
    if (maps && !two) mt2 = mt;
  }
  const bool wide = maps && wide_shape(P, O * (two ? 2 : 1));
  const Plan pl = make_plan(B, P, O, HW > 0 ? HW : 1, two, wide);
  const size_t counts_bytes = align_up((size_t)B * pl.S * pl.cnt * sizeof(int), 256);
  if (workspace_bytes < counts_bytes + 256) return DMM_ERR_WORKSPACE_TOO_SMALL;
  if (pl.n_ptiles * pl.n_otiles > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;

  IouParams kp;
  kp.item_counter = (int*)((char*)workspace + counts_bytes);
  kp.prop = prop; kp.prop_ptrs = prop_ptrs; kp.tmpl = tmpl; kp.tmpl2 = tmpl2;
  kp.prop_bs = prop_bstride; kp.tmpl_bs = tmpl_bstride; kp.tmpl2_bs = tmpl2_bstride;
  kp.n_prop = n_prop; kp.n_tmpl = n_tmpl;
  kp.P = P; kp.O = O; kp.Otot = pl.Otot; kp.HW = HW;
  kp.PT = pl.PT; kp.OT = pl.OT; kp.n_ptiles = pl.n_ptiles; kp.n_otiles = pl.n_otiles;
  kp.S = pl.S; kp.chunks_per_slab = pl.chunks_per_slab; kp.n_chunks = pl.n_chunks;
  kp.B_items = B;
  kp.ws = (int*)workspace; kp.cnt = pl.cnt;

  if (HW == 0) {  // empty masks: every count is 0 -> IoU 0/(0+1e-6) = 0; skip the streaming kernel
    DMM_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)B * pl.S * pl.cnt * sizeof(int), st));
  }
  dim3 grid(pl.S, B, pl.n_ptiles * pl.n_otiles);
  const int to = pl.OT <= 4 ? 4 : (pl.OT <= 8 ? 8 : (pl.OT <= 12 ? 12 : 16));
  const bool use_tma = maps && pl.n_ptiles == 1 && pl.n_otiles == 1;
  int rc = DMM_OK;
  if (HW == 0) {
  } else if (use_tma) {
    DMM_CUDA_TRY(cudaMemsetAsync(kp.item_counter, 0, sizeof(int), st));
    if (wide)
      rc = pl.OT <= 20 ? launch_tma<20, kWideRows, 2>(kp, mp, mt, mt2, grid, st)
         : pl.OT <= 24 ? launch_tma<24, kWideRows, 2>(kp, mp, mt, mt2, grid, st) : launch_tma<32, kWideRows, 2>(kp, mp, mt, mt2, grid, st);
    else
      rc = to == 4 ? launch_tma<4>(kp, mp, mt, mt2, grid, st) : to == 8 ? launch_tma<8>(kp, mp, mt, mt2, grid, st)
         : to == 12 ? launch_tma<12>(kp, mp, mt, mt2, grid, st) : launch_tma<16>(kp, mp, mt, mt2, grid, st);
  } else {
#define DMM_LAUNCH(V, T) mask_iou_partial_kernel<V, T><<<grid, kThreads, 0, st>>>(kp)
    if (vec) {
      if (to == 4) DMM_LAUNCH(true, 4); else if (to == 8) DMM_LAUNCH(true, 8);
      else if (to == 12) DMM_LAUNCH(true, 12); else DMM_LAUNCH(true, 16);
    } else {
      if (to == 4) DMM_LAUNCH(false, 4); else if (to == 8) DMM_LAUNCH(false, 8);
      else if (to == 12) DMM_LAUNCH(false, 12); else DMM_LAUNCH(false, 16);
    }
#undef DMM_LAUNCH
    rc = check_launch();
  }
  if (rc) return rc;

  FinParams fp;
  fp.ws = (const int*)workspace; fp.S = pl.S; fp.cnt = pl.cnt; fp.B = B; fp.P = P; fp.O = O; fp.Otot = pl.Otot;
  fp.n_prop = n_prop; fp.n_tmpl = n_tmpl; fp.iou = iou; fp.iou2 = iou2; fp.cos = cos; fp.w_cos = w_cos;
  fp.w_iou = w_iou; fp.sim = sim; fp.counts = counts;
  const long long total = (long long)B * pl.Otot * P;
  int fblocks = (int)((total + 255) / 256);
  if (fblocks > 8 * kNumSMs) fblocks = 8 * kNumSMs;
  mask_iou_finalize_kernel<<<fblocks, 256, 0, st>>>(fp);
  return check_launch();
}

extern "C" int dmm_mask_iou_pairwise(const float* prop, long long prop_bstride, const float* tmpl,
                                     long long tmpl_bstride, const float* tmpl2, long long tmpl2_bstride, int B,
                                     int P, int O, int HW, const int* n_prop, const int* n_tmpl, float* iou,
                                     float* iou2, const float* cos, float w_cos, float w_iou, float* sim,
                                     int* counts, void* workspace, size_t workspace_bytes, void* stream) {
  return run_pairwise(prop, nullptr, 0, prop_bstride, tmpl, tmpl_bstride, tmpl2, tmpl2_bstride, B, P, O, HW, n_prop,
                      n_tmpl, iou, iou2, cos, w_cos, w_iou, sim, counts, workspace, workspace_bytes, stream);
}

extern "C" int dmm_mask_iou_pairwise_ptrs(const float* const* prop_ptrs, int ptrs_aligned16, const float* tmpl,
                                          long long tmpl_bstride, const float* tmpl2, long long tmpl2_bstride, int B,
                                          int P, int O, int HW, const int* n_prop, const int* n_tmpl, float* iou,
                                          float* iou2, const float* cos, float w_cos, float w_iou, float* sim,
                                          int* counts, void* workspace, size_t workspace_bytes, void* stream) {
  if (!prop_ptrs) return DMM_ERR_INVALID_ARGUMENT;
  return run_pairwise(nullptr, prop_ptrs, ptrs_aligned16, 0, tmpl, tmpl_bstride, tmpl2, tmpl2_bstride, B, P, O, HW,
                      n_prop, n_tmpl, iou, iou2, cos, w_cos, w_iou, sim, counts, workspace, workspace_bytes, stream);
}

// =========================================================================================================
// Bit-packed masks.  bit i of word j of a row = (pixel 32*j + i) > 0.5f, zero past the row.
// The IoU only needs these bits: a producer that packs once (dmm_mask_pack_bits on the device, dmm_host_pack_masks
// on the host before the PCIe copy) moves 32x fewer bytes through K1.  Same counters, same finalize, same bits out.
// =========================================================================================================
namespace dmm {
namespace {

__global__ void __launch_bounds__(256) mask_pack_bits_kernel(const float* __restrict__ src, long long rows, int HW,
                                                            uint32_t* __restrict__ dst) {
  const int words = (HW + 31) / 32;
  const long long total = rows * words;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp0; t < total; t += nwarps) {
    const long long r = t / words;
    const int j = (int)(t - r * words);
    const int px = 32 * j + lane;
    const float v = px < HW ? src[r * HW + px] : 0.f;
    const uint32_t w = __ballot_sync(0xffffffffu, v > 0.5f);
    if (lane == 0) dst[t] = w;
  }
}

struct PackedParams {
  const uint32_t* prop; const uint32_t* tmpl; const uint32_t* tmpl2;
  long long prop_bs, tmpl_bs, tmpl2_bs;   // batch strides in words
  const int* n_prop; const int* n_tmpl;
  int P, O, Otot, words;
  int PT, OT, n_ptiles, n_otiles;
  int S, chunks_per_slab, n_chunks;       // chunk = 8 words
  int* ws; int cnt;
};

template <int TO>
__global__ void __launch_bounds__(kThreads) mask_iou_partial_packed_kernel(const PackedParams p) {
  __shared__ const uint32_t* row_ptr[kMaxRows];
  __shared__ uint32_t bits[2][2][kMaxRows + kTileO][4];
  __shared__ int red[kTileO * kMaxRows + kMaxRows];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x, b = blockIdx.y;
  const int ptile = blockIdx.z % p.n_ptiles, otile = blockIdx.z / p.n_ptiles;
  const int np = p.n_prop ? clampi(p.n_prop[b], 0, p.P) : p.P;
  const int nt = p.n_tmpl ? clampi(p.n_tmpl[b], 0, p.O) : p.O;
  const int p0 = ptile * p.PT, o0 = otile * p.OT;
  const int pcnt = min(p.PT, p.P - p0), ocnt = min(p.OT, p.Otot - o0);
  const int rows = pcnt + ocnt;
  if (tid < kMaxRows) {
    const uint32_t* ptr = nullptr;
    if (tid < pcnt) {
      if (p0 + tid < np) ptr = p.prop + (long long)b * p.prop_bs + (long long)(p0 + tid) * p.words;
    } else if (tid < rows) {
      const int t = o0 + tid - pcnt;
      if (t < p.O) {
        if (t < nt) ptr = p.tmpl + (long long)b * p.tmpl_bs + (long long)t * p.words;
      } else if (t - p.O < nt) {
        ptr = p.tmpl2 + (long long)b * p.tmpl2_bs + (long long)(t - p.O) * p.words;
      }
    }
    row_ptr[tid] = ptr;   // nullptr: padding row, contributes zero bits
  }
  for (int i = tid; i < kTileO * kMaxRows + kMaxRows; i += kThreads) red[i] = 0;
  __syncthreads();
  const int c0 = s * p.chunks_per_slab;
  const int c1 = min(c0 + p.chunks_per_slab, p.n_chunks);
  int acc[TO][2];
#pragma unroll
  for (int o = 0; o < TO; ++o) acc[o][0] = acc[o][1] = 0;
  int area0 = 0, area1 = 0;
  const int grp_b = warp >> 2, word_b = warp & 3;
  // loader role: threads 0..127 own (row = tid >> 1, group = tid & 1): four words per chunk
  const int lrow = tid >> 1, lgrp = tid & 1;
  const uint32_t* lptr = tid < 2 * kMaxRows ? row_ptr[lrow] : nullptr;
  uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
  auto fetch = [&](int c) {
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (lptr) {
      const int j = c * 8 + lgrp * 4;
      if (j + 0 < p.words) w.x = __ldg(lptr + j + 0);
      if (j + 1 < p.words) w.y = __ldg(lptr + j + 1);
      if (j + 2 < p.words) w.z = __ldg(lptr + j + 2);
      if (j + 3 < p.words) w.w = __ldg(lptr + j + 3);
    }
    return w;
  };
  if (c0 < c1) nxt = fetch(c0);
  for (int c = c0; c < c1; ++c) {
    const int buf = (c - c0) & 1;
    if (tid < 2 * kMaxRows) *reinterpret_cast<uint4*>(&bits[buf][lgrp][lrow][0]) = nxt;
    if (c + 1 < c1) nxt = fetch(c + 1);
    __syncthreads();
    const uint32_t b0 = bits[buf][grp_b][lane][word_b], b1 = bits[buf][grp_b][lane + 32][word_b];
    area0 += __popc(b0);
    area1 += __popc(b1);
#pragma unroll
    for (int o = 0; o < TO; ++o) {
      const uint32_t a = bits[buf][grp_b][pcnt + o][word_b];
      acc[o][0] += __popc(a & b0);
      acc[o][1] += __popc(a & b1);
    }
  }
#pragma unroll
  for (int o = 0; o < TO; ++o) {
    if (o < ocnt) {
      atomicAdd(&red[o * kMaxRows + lane], acc[o][0]);
      atomicAdd(&red[o * kMaxRows + lane + 32], acc[o][1]);
    }
  }
  atomicAdd(&red[kTileO * kMaxRows + lane], area0);
  atomicAdd(&red[kTileO * kMaxRows + lane + 32], area1);
  __syncthreads();
  int* out = p.ws + ((long long)b * p.S + s) * p.cnt;
  for (int i = tid; i < ocnt * pcnt; i += kThreads) {
    const int o = i / pcnt, q = i - o * pcnt;
    out[(o0 + o) * p.P + p0 + q] = red[o * kMaxRows + q];
  }
  int* area_t = out + p.Otot * p.P;
  int* area_p = area_t + p.Otot;
  if (ptile == 0)
    for (int i = tid; i < ocnt; i += kThreads) area_t[o0 + i] = red[kTileO * kMaxRows + pcnt + i];
  if (otile == 0)
    for (int i = tid; i < pcnt; i += kThreads) area_p[p0 + i] = red[kTileO * kMaxRows + i];
}

}  // namespace
}  // namespace dmm

extern "C" int dmm_mask_pack_bits(const float* masks, long long rows, int HW, uint32_t* bits, void* stream) {
  if (rows < 0 || HW < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (rows == 0 || HW == 0) return DMM_OK;
  if (!masks || !bits) return DMM_ERR_INVALID_ARGUMENT;
  const long long total = rows * ((HW + 31) / 32);
  long long blocks = (total + 7) / 8;   // 8 warps per block, one word per warp and trip
  if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
  mask_pack_bits_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(masks, rows, HW, bits);
  return check_launch();
}

extern "C" size_t dmm_mask_iou_packed_workspace_bytes(int B, int P, int O, int words, int two_template_sets) {
  // the packed kernel reuses K1's plan with one "pixel" per bit: chunk = 8 words = 256 bits
  return dmm_mask_iou_workspace_bytes(B, P, O, words > 0 ? words * 32 : 1, two_template_sets);
}

extern "C" int dmm_mask_iou_pairwise_packed(const uint32_t* prop_bits, long long prop_bstride_words,
                                            const uint32_t* tmpl_bits, long long tmpl_bstride_words,
                                            const uint32_t* tmpl2_bits, long long tmpl2_bstride_words, int B, int P,
                                            int O, int words, const int* n_prop, const int* n_tmpl, float* iou,
                                            float* iou2, const float* cos, float w_cos, float w_iou, float* sim,
                                            int* counts, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || P < 0 || O < 0 || words < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (B == 0 || P == 0 || O == 0) return DMM_OK;
  if (!workspace || (words > 0 && (!prop_bits || !tmpl_bits))) return DMM_ERR_INVALID_ARGUMENT;
  if (sim && !cos) return DMM_ERR_INVALID_ARGUMENT;
  if (iou2 && !tmpl2_bits) return DMM_ERR_INVALID_ARGUMENT;
  if (B > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;
  const int two = tmpl2_bits != nullptr;
  const Plan pl = make_plan(B, P, O, words > 0 ? words * 32 : 1, two);
  if (workspace_bytes < (size_t)B * pl.S * pl.cnt * sizeof(int)) return DMM_ERR_WORKSPACE_TOO_SMALL;
  if (pl.n_ptiles * pl.n_otiles > 65535) return DMM_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  PackedParams kp;
  kp.prop = prop_bits; kp.tmpl = tmpl_bits; kp.tmpl2 = tmpl2_bits;
  kp.prop_bs = prop_bstride_words; kp.tmpl_bs = tmpl_bstride_words; kp.tmpl2_bs = tmpl2_bstride_words;
  kp.n_prop = n_prop; kp.n_tmpl = n_tmpl; kp.P = P; kp.O = O; kp.Otot = pl.Otot; kp.words = words;
  kp.PT = pl.PT; kp.OT = pl.OT; kp.n_ptiles = pl.n_ptiles; kp.n_otiles = pl.n_otiles;
  kp.S = pl.S; kp.chunks_per_slab = pl.chunks_per_slab; kp.n_chunks = pl.n_chunks;
  kp.ws = (int*)workspace; kp.cnt = pl.cnt;
  if (words == 0) {
    DMM_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)B * pl.S * pl.cnt * sizeof(int), st));
  } else {
    dim3 grid(pl.S, B, pl.n_ptiles * pl.n_otiles);
    const int to = pl.OT <= 4 ? 4 : (pl.OT <= 8 ? 8 : (pl.OT <= 12 ? 12 : 16));
    if (to == 4) mask_iou_partial_packed_kernel<4><<<grid, kThreads, 0, st>>>(kp);
    else if (to == 8) mask_iou_partial_packed_kernel<8><<<grid, kThreads, 0, st>>>(kp);
    else if (to == 12) mask_iou_partial_packed_kernel<12><<<grid, kThreads, 0, st>>>(kp);
    else mask_iou_partial_packed_kernel<16><<<grid, kThreads, 0, st>>>(kp);
    int rc = check_launch();
    if (rc) return rc;
  }
  FinParams fp;
  fp.ws = (const int*)workspace; fp.S = pl.S; fp.cnt = pl.cnt; fp.B = B; fp.P = P; fp.O = O; fp.Otot = pl.Otot;
  fp.n_prop = n_prop; fp.n_tmpl = n_tmpl; fp.iou = iou; fp.iou2 = iou2; fp.cos = cos; fp.w_cos = w_cos;
  fp.w_iou = w_iou; fp.sim = sim; fp.counts = counts;
  const long long total = (long long)B * pl.Otot * P;
  int fblocks = (int)((total + 255) / 256);
  if (fblocks > 8 * kNumSMs) fblocks = 8 * kNumSMs;
  mask_iou_finalize_kernel<<<fblocks, 256, 0, st>>>(fp);
  return check_launch();
}

extern "C" size_t dmm_mask_iou_rowwise_workspace_bytes(int N, int M) {
  return dmm_mask_iou_workspace_bytes(N, 1, 1, M, 0);
}

extern "C" int dmm_mask_iou_rowwise(const float* a, const float* b, int N, int M, float* iou, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  // row i of a against row i of b == N problems with one proposal and one template each
  if (N < 0 || M < 0) return DMM_ERR_INVALID_ARGUMENT;
  if (N == 0) return DMM_OK;
  if (!iou) return DMM_ERR_INVALID_ARGUMENT;
  int done = 0;
  while (done < N) {  // grid.y limit
    const int nb = N - done < 65535 ? N - done : 65535;
    int rc = dmm_mask_iou_pairwise(b + (long long)done * M, M, a + (long long)done * M, M, nullptr, 0, nb, 1, 1, M,
                                   nullptr, nullptr, iou + done, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr,
                                   workspace, workspace_bytes, stream);
    if (rc) return rc;
    done += nb;
  }
  return DMM_OK;
}
