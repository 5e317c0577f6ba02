// K5-TC -- proposal-feature pooling as a per-frame dense contraction on the 5th-generation tensor cores, TMA-fed.
//
// Reference: dmm/modules/feature_extractor.py:11-52 (legacy ROIAlign 14x14 / sampling 2 on 4 levels + spatial mean over
// maskrcnn_benchmark's un-vendored Pooler; stand-in oracle torchvision roi_align(aligned=False)).  The mean of the
// bilinear samples is linear in the feature map and separable:
//     out[r, l*C + c] = sum_y wy[r,l,y] * ( sum_x wx[r,l,x] * F_l[n_r, c, y, x] )
// The SIMT kernel (roi_mean_pool.cu) gathers every ROI's window per channel from L2: a level-0 map is re-read by every one
// of the frame's ~50 ROIs.  Here the ROIs of one frame form the N dimension of a GEMM, so every feature byte is read from
// HBM ONCE per frame:  for each feature row y,   D_y[c, r] = sum_x F[c, y, x] * wx[r, x]     (M = 128 channels,
// N = 64 ROI slots, K = W_l) on tcgen05 (3xTF32 -> fp32-grade products), and the epilogue folds the rows with fp32 FMAs:
// acc[c, r] += wy[r, y] * D_y[c, r].  Per row only W_l/8 K-steps x 3 MMAs are chained in one accumulator, so the
// tensor core's truncating accumulation (see cosine_tc.cu) never builds up, and the B operand (wx, split into TF32 hi/lo,
// K-major SWIZZLE_128B) is the same for every row of the frame: built once per work item, not per chunk.
// Levels whose row pitch TMA cannot address (W_l % 4 != 0, e.g. the 8x14 level) but that are small (H*W <= 128) run in
// "linear" mode: the whole map is ONE row of H*W elements and B holds the full outer product wy (x) wx.
//
// Three launches, no host synchronisation:
//   tc_group_kernel    buckets the ROIs by frame into groups of <= 64 (any order of the input rows), zeroes counters;
//   tc_weights_kernel  per (group, level): the 1-D weight vectors of its ROIs (same sample arithmetic and summation
//                      order as roi_mean_pool.cu's axis_weight), transposed wy for the epilogue, the union window;
//   roi_pool_tc_kernel persistent, warp-specialised, one CTA per SM; work item = (group, level, x segment, band of
//                      rows), claimed from an atomic counter, big items first:
//        warp 16 (1 lane) TMA: one cp.async.bulk.tensor.3d box [128 ch x 1 row x 32 px] (SWIZZLE_128B, zero fill past
//                         W) per chunk into a 4-deep ring; only rows / x chunks inside the union window of the group's
//                         ROIs are fetched;
//        warps 0-7        split pass (two groups on alternate chunks): fp32 -> TF32 hi / lo, written to TENSOR MEMORY
//                         (tcgen05.st; channel = TMEM lane) as the A operand;
//        warps 18-19      build the next item's B tiles (wx hi / lo) in the other half of a double buffer;
//        warp 17 (1 lane) 12 x tcgen05.mma.kind::tf32 M128 N64 K8 per chunk, A from TMEM, B from shared memory;
//                         tcgen05.commit publishes the row's accumulator (two accumulator buffers);
//        warps 8-15       epilogue: tcgen05.ld of the row's accumulator, acc += wy * D_y; at the end of the item the
//                         [ROI x channel] partial goes to `out` (single item) or to the workspace, where the LAST item
//                         of the (group, level) adds the partials in item order -- deterministic, no float atomics.
// Band sizes depend on the feature-map shapes and the number of frames only, never on the ROIs: the same box gives the
// same bits whatever table it arrives in (the lazy pipeline's bit-identity test relies on it).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace dmm {
namespace {

using namespace tc;

constexpr int kC = 128;                         // channels = MMA M (the reference's neck: 128 per level)
constexpr int kNR = 64;                         // ROI slots per group = MMA N
constexpr int kKC = 32;                         // floats per K chunk (one 128-byte swizzle row)
#ifndef K5_STAGES
#define K5_STAGES 4
#endif
// Raw fp32 ring (what the TMA keeps in flight: 4 x 16 KB; measured at 64 frames x 50 ROIs: 2 stages 150 us, 4 stages 136 us --
// the ring is not the limiter, the converter -> MMA -> epilogue hand-offs are).  MUST BE EVEN:
// the two converter groups take alternate chunks, and a waiter tests the PARITY of an mbarrier phase.  With an odd ring
// the groups alternate on every stage, so a group's consecutive visits to a stage are two phases apart -- same parity --
// and its wait for chunk g+2*kStages passes as soon as phase g has completed, i.e. possibly BEFORE chunk g+kStages (the
// other group's) has even landed: TMA completions are not ordered.  Seen once per ~500 calls on 8 GPUs under NVLink
// traffic (5 stages): stale data converted, arrivals on the wrong phase, pipeline stall.  With an even ring a stage always
// belongs to the same group and consecutive visits are consecutive phases.
constexpr int kStages = K5_STAGES;
#ifndef K5_ALLOW_ODD_RING   // scripts/k5_stress_nccl.py builds the faulty odd ring on purpose to show the race
static_assert(kStages % 2 == 0, "raw ring depth must be even (see above)");
#endif
constexpr int kOpStages = 4;                    // A operand ring in TMEM
constexpr int kMetaRing = 16;                   // per-chunk metadata (producer leads the MMA warp by < 10 chunks)
constexpr int kItemRing = 4;
constexpr int kMaxXC = 4;                       // x chunks per segment (128 px)
constexpr int kSegW = kMaxXC * kKC;
constexpr uint32_t kRawBytes = kC * kKC * 4;    // 16 KB
constexpr uint32_t kBTile = kNR * kKC * 4;      // 8 KB (one of hi / lo)
constexpr uint32_t kBBuf = kMaxXC * 2 * kBTile; // 64 KB: 4 x chunks x (hi + lo)
constexpr int kMaxBandRows = 16;                // rows of one work item: bounds the wy tile the epilogue reads from shared memory
constexpr uint32_t kWyTile = kMaxBandRows * kNR * 4;   // 4 KB
constexpr int kWyTiles = 4;                     // the epilogue may still read the tile of item i-3 while item i's is written
constexpr size_t kDynSmem = (size_t)kStages * kRawBytes + 2 * (size_t)kBBuf + kWyTiles * (size_t)kWyTile + 1024;   // 225 KB
constexpr int kConvWarps = 8, kEpiWarp0 = 8, kEpiWarps = 8, kTmaWarp = 16, kMmaWarp = 17, kLoadWarp = 18;
constexpr int kThreads = 19 * 32;
constexpr uint32_t kAccCols = kNR;              // one accumulator: 64 fp32 columns
constexpr uint32_t kOpCol0 = 2 * kAccCols;      // A ring behind the two accumulators
constexpr uint32_t kOpCols = 2 * kKC;           // hi + lo
constexpr uint32_t kTmemCols = 512;
constexpr int kRes = 14, kSamp = 2, kNS = kRes * kSamp;   // feature_extractor.py:14-15
constexpr int kMaxFrames = 4096;                // frame histogram lives in shared memory

enum { MODE_NONE = 0, MODE_ROW = 1, MODE_LIN = 2 };

struct LevelInfo {
  int H, W;            // the feature map
  int mode;
  int Hv, Wv;          // virtual rows x row length seen by the GEMM (ROW: H x W; LIN: 1 x H*W)
  int nseg;            // x segments of <= 128 elements
  int rows_per_band, nband;
  int items;           // nseg * nband = work items (and partials) of one (group, level)
  int item0;           // first item of this level within a group
  int wy_off, wx_off;  // float offsets of wyT [H][64] and wx [64][Wp] inside a group's table
  int Wp;              // W rounded up to 4
  int nxc;             // 32-element chunks of a virtual row (all segments)
  int bt_off;          // float offset of this level's B tiles [nxc][hi|lo][64][32] inside a group's tile table
};

struct TcPool {
  LevelInfo lv[4];
  int N, R, Gb, items_per_group, wtab_stride;
  const float* rois;
  float* out;
  int* hdr;        // [0] = G (groups), [1] = item counter
  int* grp_frame;  // [Gb]
  int* grp_start;  // [Gb] first entry in perm
  int* grp_count;  // [Gb]
  int* perm;       // [R] ROI ids bucketed by frame
  int* pair_ctr;   // [Gb*4]
  int* ranges;     // [Gb*4][4] union window (ylo, yhi, xlo, xhi) in virtual coordinates; ylo > yhi: empty
  float* wtab;     // [Gb][wtab_stride]
  float* btab;     // [Gb][btab_stride]  B operand tiles, TF32 hi / lo, already in the K-major SWIZZLE_128B layout
  int btab_stride;
  float* partial;  // [Gb*items_per_group][64][128]
  long long* trace;  // debug builds (DMM_TC_DEBUG + DMM_K5_TRACE_FILE): (event, clock64) pairs of CTA 0
  int tables_only;   // backward: weight tables (wyT, wx) for EVERY level, no B tiles / ranges
  int C;             // channels (the contraction kernel needs 128; the backward gather takes any)
};

struct Item {      // what the producer publishes per work item (shared memory ring)
  int g, l, n, nroi, start;
  int ya, yb;      // rows [ya, yb) actually streamed
  int jlo, jhi;    // B tiles [jlo, jhi] of the segment actually used (relative to the segment)
  int x0;          // first element of the segment
  int bseq;        // sequence number among non-empty items (B double buffer), -1: nothing streamed
  int slot;        // partial slot, pair index
  int stop;
};

// ---------------------------------------------------------------------------------------------------------------------
// bucketing
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) tc_group_kernel(const TcPool p) {
  __shared__ int cnt[kMaxFrames], roi0[kMaxFrames];
  __shared__ int warp_r[32], warp_g[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < p.N; i += 1024) cnt[i] = 0;
  for (int i = tid; i < 4 * p.Gb; i += 1024) p.pair_ctr[i] = 0;
  __syncthreads();
  for (int i = tid; i < p.R; i += 1024) {
    const int n = (int)p.rois[(long long)i * 5];
    if (n >= 0 && n < p.N) atomicAdd(&cnt[n], 1);
    else if (p.out) {                                       // ROIs of no frame pool to zero (as the SIMT kernel)
      float* o = p.out + (long long)i * 4 * kC;
      for (int c = 0; c < 4 * kC; ++c) o[c] = 0.f;
    }
  }
  __syncthreads();
  // exclusive scan of (ROI count, group count) over the frames: 4 frames per thread, warp scan, scan of warp totals
  int c4[4], lr = 0, lg = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = tid * 4 + k;
    c4[k] = f < p.N ? cnt[f] : 0;
    lr += c4[k];
    lg += (c4[k] + kNR - 1) / kNR;
  }
  int sr = lr, sg = lg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, sr, o), b = __shfl_up_sync(0xffffffffu, sg, o);
    if (lane >= o) { sr += a; sg += b; }
  }
  if (lane == 31) { warp_r[warp] = sr; warp_g[warp] = sg; }
  __syncthreads();
  if (warp == 0) {
    int a = warp_r[lane], b = warp_g[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += x; b += y; }
    }
    warp_r[lane] = a; warp_g[lane] = b;                      // inclusive
  }
  __syncthreads();
  int er = sr - lr + (warp ? warp_r[warp - 1] : 0), eg = sg - lg + (warp ? warp_g[warp - 1] : 0);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = tid * 4 + k;
    if (f < p.N) {
      roi0[f] = er;
      const int ng = (c4[k] + kNR - 1) / kNR;
      for (int q = 0; q < ng; ++q) {
        p.grp_frame[eg + q] = f;
        p.grp_start[eg + q] = er + q * kNR;
        p.grp_count[eg + q] = min(kNR, c4[k] - q * kNR);
      }
      er += c4[k]; eg += ng;
    }
  }
  if (tid == 1023) { p.hdr[0] = warp_g[31]; p.hdr[1] = 0; }
  __syncthreads();
  // perm: the ROIs of a frame in INPUT ORDER (stable) -- the slot of a ROI inside its group fixes the order in which the
  // backward gather adds its contribution, so it must not depend on the arrival order of atomics.
  // Fast path: rows already grouped by ascending frame id with no invalid row (what convert_to_roi_format produces).
  __shared__ int s_unsorted;
  if (tid == 0) s_unsorted = 0;
  __syncthreads();
  for (int i = tid; i < p.R; i += 1024) {
    const int n = (int)p.rois[(long long)i * 5];
    const int prev = i ? (int)p.rois[(long long)(i - 1) * 5] : 0;
    if (n < 0 || n >= p.N || n < prev) s_unsorted = 1;
  }
  for (int i = tid; i < p.N; i += 1024) cnt[i] = 0;          // reuse as cursors
  __syncthreads();
  if (!s_unsorted) {
    for (int i = tid; i < p.R; i += 1024) p.perm[i] = i;
    return;
  }
  // general path: stable counting sort, 1024 rows at a time; inside a warp the rows of one frame are ranked with
  // match.any, the 32 warps then take turns on the per-frame cursors
  for (int c0 = 0; c0 < p.R; c0 += 1024) {
    const int i = c0 + tid;
    const int n = i < p.R ? (int)p.rois[(long long)i * 5] : -1;
    const bool ok = n >= 0 && n < p.N;
    const unsigned peers = __match_any_sync(0xffffffffu, ok ? n : -1 - lane);
    const int rank = __popc(peers & ((1u << lane) - 1u)), leader = __ffs(peers) - 1;
    int base = 0;
    for (int w = 0; w < 32; ++w) {
      if (warp == w && ok && lane == leader) { base = cnt[n]; cnt[n] = base + __popc(peers); }
      __syncthreads();
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) p.perm[roi0[n] + base + rank] = i;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// weights: wy / wx of every ROI of a (group, level); same samples, same order of additions as axis_weight() in
// roi_mean_pool.cu, so both kernels contract with bit-identical weights.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_weights_kernel(const TcPool p) {
  extern __shared__ float wsm[];                              // wy [64][H] | wx [64][W]
  __shared__ int rng[4];
  const int g = blockIdx.x, l = blockIdx.y, tid = threadIdx.x;
  if (g >= p.hdr[0]) return;
  const LevelInfo lv = p.lv[l];
  if (lv.mode == MODE_NONE && !p.tables_only) return;
  const int H = lv.H, W = lv.W;
  float* wy = wsm;
  float* wx = wsm + kNR * H;
  for (int i = tid; i < kNR * (H + W); i += 256) wsm[i] = 0.f;
  if (tid == 0) { rng[0] = H; rng[1] = -1; rng[2] = W; rng[3] = -1; }
  __syncthreads();
  const int nroi = p.grp_count[g], start = p.grp_start[g];
  if (tid < 2 * kNR) {
    const int r = tid >> 1, axis = tid & 1;                   // axis 0: y, 1: x
    if (r < nroi) {
      const float* roi = p.rois + (long long)p.perm[start + r] * 5;
      const float lo = axis ? roi[1] : roi[2], hi = axis ? roi[3] : roi[4];
      const int size = axis ? W : H;
      float* w = (axis ? wx : wy) + r * size;
      const float scale = 0.25f / (float)(1 << l);            // feature_extractor.py:13
      const float st = lo * scale;
      const float len = fmaxf(hi * scale - st, 1.f);
      const float bin = len / (float)kRes;
      int imin = size, imax = -1;
      for (int s = 0; s < kNS; ++s) {
        const int bidx = s / kSamp, sidx = s - bidx * kSamp;
        float t = st + (float)bidx * bin + ((float)sidx + 0.5f) * bin / (float)kSamp;
        if (t < -1.f || t > (float)size) continue;
        if (t <= 0.f) t = 0.f;
        int a = (int)t, b;
        if (a >= size - 1) { a = b = size - 1; t = (float)a; } else b = a + 1;
        const float fr = t - (float)a;
        w[a] += 1.f - fr;
        w[b] += fr;
        imin = min(imin, a); imax = max(imax, b);
      }
      for (int i = max(imin, 0); i <= imax; ++i) w[i] *= (1.f / (float)kNS);
      if (imax >= 0) { atomicMin(&rng[axis ? 2 : 0], imin); atomicMax(&rng[axis ? 3 : 1], imax); }
    }
  }
  __syncthreads();
  float* tab = p.wtab + (long long)g * p.wtab_stride;
  float* wyT = tab + lv.wy_off;                               // [H][64]
  float* wxg = tab + lv.wx_off;                               // [64][Wp]
  for (int i = tid; i < H * kNR; i += 256) { const int y = i / kNR, r = i % kNR; wyT[i] = wy[r * H + y]; }
  for (int i = tid; i < kNR * lv.Wp; i += 256) { const int r = i / lv.Wp, x = i % lv.Wp; wxg[i] = x < W ? wx[r * W + x] : 0.f; }
  if (p.tables_only) return;
  // The MMA's B operand for every 32-element chunk of the (virtual) row, exactly as shared memory wants it: 64 rows (ROI
  // slots) x 128 bytes, 16-byte chunk c of row r at ((c ^ (r & 7)) << 4), TF32 hi tile then lo tile.  Built once per
  // (group, level) here; the contraction kernel only bulk-copies it (every band of the level reuses it from L2).
  {
    float* bt = p.btab + (long long)g * p.btab_stride + lv.bt_off;
    const int n16 = lv.nxc * kNR * 8;                         // 16-byte pieces per hi (or lo) set of tiles
    for (int q = tid; q < n16; q += 256) {
      const int j = q / (kNR * 8), r = (q >> 3) % kNR, c = q & 7;
      float w4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = j * kKC + c * 4 + e;                    // element of the virtual row
        float w = 0.f;
        if (lv.mode == MODE_ROW) { if (k < W) w = wx[r * W + k]; }
        else if (k < H * W) { const int y = k / W, x = k - y * W; w = __fmul_rn(wy[r * H + y], wx[r * W + x]); }
        w4[e] = w;
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = __float_as_uint(w4[e]) & 0xffffe000u;
        lo[e] = __float_as_uint(__fsub_rn(w4[e], __uint_as_float(hi[e]))) & 0xffffe000u;
      }
      float* dst = bt + (long long)j * (2 * kNR * kKC) + r * kKC + ((c ^ (r & 7)) << 2);
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + kNR * kKC) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  if (tid == 0) {
    int* o = p.ranges + ((long long)g * 4 + l) * 4;
    const bool empty = rng[1] < 0 || rng[3] < 0;
    if (lv.mode == MODE_ROW) {
      o[0] = empty ? 1 : rng[0]; o[1] = empty ? 0 : rng[1]; o[2] = rng[2]; o[3] = rng[3];
    } else {                                                  // LIN: one virtual row of H*W elements
      o[0] = empty ? 1 : 0; o[1] = 0;
      o[2] = empty ? 0 : rng[0] * W; o[3] = empty ? 0 : rng[1] * W + W - 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the contraction
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_item(const TcPool& p, int t, int G, Item& it) {
  // claim order: every level-0 item of every group, then level 1, ... (big items first)
  int l = 0, base = 0;
  bool found = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int n = G * p.lv[k].items;
    if (!found) {
      if (t < base + n || k == 3) { l = k; found = true; } else base += n;
    }
  }
  const LevelInfo& lv = p.lv[l];
  const int rel = t - base, g = rel / lv.items, rem = rel % lv.items;
  const int seg = rem / lv.nband, band = rem % lv.nband;
  it.g = g; it.l = l;
  it.n = p.grp_frame[g]; it.nroi = p.grp_count[g]; it.start = p.grp_start[g];
  const int* rg = p.ranges + ((long long)g * 4 + l) * 4;
  const int ylo = rg[0], yhi = rg[1], xlo = rg[2], xhi = rg[3];
  it.x0 = seg * kSegW;
  int ya = max(band * lv.rows_per_band, ylo), yb = min(min((band + 1) * lv.rows_per_band, lv.Hv), yhi + 1);
  const int sx0 = max(xlo, it.x0), sx1 = min(xhi, min(it.x0 + kSegW, lv.Wv) - 1);   // union window inside this segment
  if (sx1 < sx0) yb = ya;                                                          // nothing of the window in this segment
  it.ya = ya; it.yb = max(yb, ya);
  it.jlo = sx1 >= sx0 ? (sx0 - it.x0) / kKC : 0;
  it.jhi = sx1 >= sx0 ? (sx1 - it.x0) / kKC : 0;
  it.slot = g * p.items_per_group + lv.item0 + rem;
  it.stop = 0;
}

__global__ void __launch_bounds__(kThreads, 1) roi_pool_tc_kernel(const TcPool p, const __grid_constant__ CUtensorMap map0,
                                                                  const __grid_constant__ CUtensorMap map1,
                                                                  const __grid_constant__ CUtensorMap map2,
                                                                  const __grid_constant__ CUtensorMap map3) {
  extern __shared__ uint8_t dyn_raw[];
  __shared__ uint64_t raw_full[kStages], raw_empty[kStages], op_full[kOpStages], op_empty[kOpStages];
  __shared__ uint64_t acc_full[2], acc_empty[2], b_full[2], b_empty[2], item_full[kItemRing], item_empty[kItemRing];
  __shared__ uint64_t wy_full[kWyTiles];
  __shared__ uint32_t s_meta[kMetaRing];
  __shared__ Item s_item[kItemRing];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_last;
#ifdef DMM_TC_DEBUG
  __shared__ volatile int s_prog[8];       // producer chunk / item, converter groups, MMA chunk, epilogue row / item, builder item
#define PROG(i, v) do { if ((threadIdx.x & 31) == 0) s_prog[i] = (int)(v); } while (0)
#define mbar_wait(bar, par, id) do { \
    const long long _t0 = clock64(); \
    while (!mbar_try(bar, par)) { \
      if (clock64() - _t0 > 300000000LL) { \
        if ((threadIdx.x & 31) == 0) printf("timeout blk %d warp %d id %d par %u | prod gch %d item %d conv %d %d mma %d epi rc %d item %d build %d\n", \
          blockIdx.x, threadIdx.x >> 5, id, (unsigned)(par), s_prog[0], s_prog[1], s_prog[2], s_prog[3], s_prog[4], s_prog[5], s_prog[6], s_prog[7]); \
        __trap(); } } } while (0)
  __shared__ int s_ntrace;
  if (threadIdx.x == 0) s_ntrace = 0;
#define TRACE(ev) do { if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) { const int _i = atomicAdd(&s_ntrace, 1); \
    if (_i < 2048) { p.trace[2 * _i] = (ev); p.trace[2 * _i + 1] = clock64(); } } } while (0)
#else
#define PROG(i, v) do { } while (0)
#define TRACE(ev) do { } while (0)
#endif
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t smem_base = smem_u32(dyn_raw);
  uint8_t* const smem_gen = dyn_raw;
  const uint32_t stage0 = (smem_base + 1023u) & ~1023u;
  const uint32_t bbuf0 = stage0 + kStages * kRawBytes;
  const uint32_t wy0 = bbuf0 + 2 * kBBuf;                    // wy tiles [rows of the item][64 ROI slots], ring of 4

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&raw_full[s]), 1); mbar_init(smem_u32(&raw_empty[s]), kConvWarps / 2); }
    for (int l = 0; l < kOpStages; ++l) { mbar_init(smem_u32(&op_full[l]), kConvWarps / 2); mbar_init(smem_u32(&op_empty[l]), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&acc_full[a]), 1); mbar_init(smem_u32(&acc_empty[a]), kEpiWarps);
      mbar_init(smem_u32(&b_full[a]), 1); mbar_init(smem_u32(&b_empty[a]), 1);
    }
    for (int i = 0; i < kItemRing; ++i) { mbar_init(smem_u32(&item_full[i]), 1); mbar_init(smem_u32(&item_empty[i]), kEpiWarps + 1); }
    for (int i = 0; i < kWyTiles; ++i) mbar_init(smem_u32(&wy_full[i]), 1);
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

  if (warp == kTmaWarp) {
    // ===== producer: claims items, publishes them, streams their chunks =====
    const int G = p.hdr[0];
    const int total = G * p.items_per_group;
    uint32_t gch = 0, iseq = 0;
    int bseq = 0;
    // publish = hand the item to the loader / epilogue warps through the item ring.  Items are published ONE AHEAD of the
    // item being streamed, so that the loader fetches the next item's B tiles while this item's chunks flow.
    auto publish = [&](Item& it) {
      it.bseq = it.yb > it.ya ? bseq++ : -1;
      const uint32_t islot = iseq % kItemRing;
      if (iseq >= kItemRing) mbar_wait(smem_u32(&item_empty[islot]), ((iseq / kItemRing) - 1u) & 1u, 1);
      if (lane == 0) { s_item[islot] = it; mbar_arrive(smem_u32(&item_full[islot])); }
      __syncwarp();
      ++iseq;
    };
    int t = 0;
    if (lane == 0) t = atomicAdd(&p.hdr[1], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    Item cur;
    if (t < total) { decode_item(p, t, G, cur); publish(cur); }
    while (t < total) {
      int tn = total;
      Item nxt;
      bool claimed = false;
      // The next item is claimed, decoded (dependent global loads, ~1.3 us) and published right after this item's first
      // kStages chunks have been issued: the ring is full then and the converters stay fed while the producer is away.
      auto claim_next = [&]() {
        TRACE(1);
        if (lane == 0) tn = atomicAdd(&p.hdr[1], 1);
        tn = __shfl_sync(0xffffffffu, tn, 0);
        if (tn < total) { decode_item(p, tn, G, nxt); publish(nxt); }
        claimed = true;
        TRACE(2);
      };
      const CUtensorMap* map = cur.l == 0 ? &map0 : (cur.l == 1 ? &map1 : (cur.l == 2 ? &map2 : &map3));
      const int zrow = cur.n * kC;
      int issued = 0;
      for (int y = cur.ya; y < cur.yb; ++y) {
        for (int j = cur.jlo; j <= cur.jhi; ++j, ++gch) {
          if (issued++ == kStages) claim_next();
          PROG(0, gch); PROG(1, iseq);
          const uint32_t s = gch % kStages;
          if (gch >= kStages) mbar_wait(smem_u32(&raw_empty[s]), ((gch / kStages) - 1u) & 1u, 2);
          if (elect_one()) {
            const uint32_t first_item = (y == cur.ya && j == cur.jlo), last_item = (y == cur.yb - 1 && j == cur.jhi);
            s_meta[gch % kMetaRing] = ((j == cur.jlo) << 1) | ((j == cur.jhi) << 2) | ((uint32_t)j << 3) | ((uint32_t)(cur.bseq & 1) << 5) |
                                      (first_item << 6) | (last_item << 7) | ((uint32_t)((cur.nroi + 15) / 16 - 1) << 8);
            const uint32_t bar = smem_u32(&raw_full[s]);
            mbar_expect_tx(bar, kRawBytes);
            tma_load_3d(stage0 + s * kRawBytes, map, cur.x0 + j * kKC, y, zrow, bar);
          }
          __syncwarp();
        }
      }
      TRACE(3);
      if (!claimed) claim_next();
      t = tn; cur = nxt;
    }
    // sentinels: one stop item for the epilogue / builder warps, two stop chunks (one per converter group; the first also
    // stops the MMA warp)
    {
      const uint32_t islot = iseq % kItemRing;
      if (iseq >= kItemRing) mbar_wait(smem_u32(&item_empty[islot]), ((iseq / kItemRing) - 1u) & 1u, 1);
      if (lane == 0) { s_item[islot].stop = 1; mbar_arrive(smem_u32(&item_full[islot])); }
      __syncwarp();
      for (int k = 0; k < 2; ++k, ++gch) {
        const uint32_t s = gch % kStages;
        if (gch >= kStages) mbar_wait(smem_u32(&raw_empty[s]), ((gch / kStages) - 1u) & 1u, 2);
        if (lane == 0) { s_meta[gch % kMetaRing] = 1u; mbar_arrive(smem_u32(&raw_full[s])); }
        __syncwarp();
      }
    }
  } else if (warp < kConvWarps) {
    // ===== split pass: fp32 feature chunk -> TF32 hi / lo in tensor memory (the A operand; TMEM lane = channel) =====
    const uint32_t grp = (uint32_t)(warp >> 2);
    const int row = (warp & 3) * 32 + lane;
    const uint32_t row_off = (uint32_t)row * 128u, sw = (uint32_t)(row & 7);
    for (uint32_t g = grp;; g += 2) {
      if ((warp & 3) == 0) PROG(2 + grp, g);
      const uint32_t s = g % kStages, l = g % kOpStages;
      mbar_wait(smem_u32(&raw_full[s]), (g / kStages) & 1u, 3);
      if (s_meta[g % kMetaRing] & 1u) {                       // stop: pass it on to the MMA warp -- behind the same wait as a real
        // chunk: arriving while the MMA warp still waits for chunk g-4 on this barrier would complete two phases in a row
        // and the waiter (which tests the phase parity) would never see its own
        if (g >= kOpStages) mbar_wait(smem_u32(&op_empty[l]), ((g / kOpStages) - 1u) & 1u, 4);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&op_full[l]));
        break;
      }
      const uint32_t raw = stage0 + s * kRawBytes + row_off;
      uint4 x[8];
#pragma unroll
      for (uint32_t c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const uint4*>(smem_gen + (raw + ((c ^ sw) << 4) - smem_base));
      if (g >= kOpStages) mbar_wait(smem_u32(&op_empty[l]), ((g / kOpStages) - 1u) & 1u, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kOpCol0 + l * kOpCols;
#pragma unroll
      for (uint32_t h = 0; h < 2; ++h) {
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
        tmem_st16(taddr + 16u * h, hi);
        tmem_st16(taddr + kKC + 16u * h, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&raw_empty[s]));
        mbar_arrive(smem_u32(&op_full[l]));
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    uint32_t rc = 0, nb = 0;                                  // rows finished, non-empty items started
    for (uint32_t g = 0;; ++g) {
      PROG(4, g);
      const uint32_t l = g % kOpStages;
      mbar_wait(smem_u32(&op_full[l]), (g / kOpStages) & 1u, 5);
      const uint32_t meta = s_meta[g % kMetaRing];
      if (meta & 1u) break;
      const uint32_t first = (meta >> 1) & 1u, last = (meta >> 2) & 1u, j = (meta >> 3) & 3u, bb = (meta >> 5) & 1u;
      if (meta & 64u) { TRACE(7); mbar_wait(smem_u32(&b_full[bb]), (nb >> 1) & 1u, 6); ++nb; TRACE(8); }      // this item's B tiles are built
      const uint32_t ab = rc & 1u;
      if (first && rc >= 2) mbar_wait(smem_u32(&acc_empty[ab]), ((rc >> 1) - 1u) & 1u, 7);   // epilogue has read row rc-2
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * kAccCols;
      const uint32_t a_hi = tmem_base + kOpCol0 + l * kOpCols, a_lo = a_hi + kKC;
      const uint32_t b_hi = bbuf0 + bb * kBBuf + j * 2 * kBTile, b_lo = b_hi + kBTile;
      const uint64_t db_hi = umma_desc(b_hi), db_lo = umma_desc(b_lo);
      const uint32_t idesc = idesc_tf32(kC, 16 * (int)(((meta >> 8) & 3u) + 1u));   // N = ROI slots in use, rounded up to 16
      if (elect_one()) {
#pragma unroll
        for (uint32_t k = 0; k < kKC / 8; ++k) {
          tc_mma_tf32_ts(d_tmem, a_lo + 8 * k, db_hi + 2 * k, idesc, !(first && k == 0));
          tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, db_lo + 2 * k, idesc, 1u);
          tc_mma_tf32_ts(d_tmem, a_hi + 8 * k, db_hi + 2 * k, idesc, 1u);
        }
        tc_commit(smem_u32(&op_empty[l]));
        if (last) tc_commit(smem_u32(&acc_full[ab]));
        if (meta & 128u) tc_commit(smem_u32(&b_empty[bb]));   // last chunk of the item: its B buffer may be rebuilt
      }
      __syncwarp();
      if (meta & 128u) TRACE(9);
      if (last) ++rc;
    }
  } else if (warp == kLoadWarp) {
    // ===== loader: the item's B tiles (built by tc_weights_kernel, already swizzled) and wy rows, two bulk copies =====
    for (uint32_t iseq = 0;; ++iseq) {
      const uint32_t islot = iseq % kItemRing;
      mbar_wait(smem_u32(&item_full[islot]), (iseq / kItemRing) & 1u, 8);
      const Item it = s_item[islot];
      PROG(7, iseq);
      if (it.stop) break;
      if (it.bseq >= 0) {
        const uint32_t bb = (uint32_t)it.bseq & 1u;
        TRACE(4);
        if (it.bseq >= 2) mbar_wait(smem_u32(&b_empty[bb]), (((uint32_t)it.bseq >> 1) - 1u) & 1u, 9);   // MMAs of item bseq-2 done
        TRACE(5);
        if (elect_one()) {
          const LevelInfo& lv = p.lv[it.l];
          const int j0 = it.x0 / kKC;                          // first chunk of the segment within the virtual row
          const float* src = p.btab + (long long)it.g * p.btab_stride + lv.bt_off + (long long)(j0 + it.jlo) * (2 * kNR * kKC);
          const uint32_t bytes = (uint32_t)(it.jhi - it.jlo + 1) * 2u * kBTile;
          const uint32_t bar = smem_u32(&b_full[bb]);
          mbar_expect_tx(bar, bytes);
          bulk_load(bbuf0 + bb * kBBuf + (uint32_t)it.jlo * 2u * kBTile, src, bytes, bar);
          // wy rows of the item for the epilogue.  Ring of 4 tiles: the loader is at most kItemRing items ahead of the
          // epilogue (item ring), so tile bseq % 4 is free when item bseq is loaded.
          const uint32_t wbar = smem_u32(&wy_full[(uint32_t)it.bseq % kWyTiles]);
          if (lv.mode == MODE_ROW) {
            const uint32_t wbytes = (uint32_t)(it.yb - it.ya) * kNR * 4u;
            mbar_expect_tx(wbar, wbytes);
            bulk_load(wy0 + ((uint32_t)it.bseq % kWyTiles) * kWyTile, p.wtab + (long long)it.g * p.wtab_stride + lv.wy_off + (long long)it.ya * kNR,
                      wbytes, wbar);
          } else {
            mbar_arrive(wbar);
          }
        }
        __syncwarp();
        TRACE(6);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&item_empty[islot]));
    }
  } else {
    // ===== epilogue: acc[c, r] += wy[r, y] * D_y[c, r]; write / reduce the item's partial =====
    const int ew = warp - kEpiWarp0;
    const int q4 = ew & 3, half = ew >> 2;
    const int c = q4 * 32 + lane;                             // channel = TMEM lane
    const int et = ew * 32 + lane;                            // 0..255
    uint32_t rc = 0;
    for (uint32_t iseq = 0;; ++iseq) {
      const uint32_t islot = iseq % kItemRing;
      mbar_wait(smem_u32(&item_full[islot]), (iseq / kItemRing) & 1u, 8);
      const Item it = s_item[islot];
      if (it.stop) break;
      const LevelInfo& lv = p.lv[it.l];
      if (ew == 0) TRACE(10);
      const float* wy_s = reinterpret_cast<const float*>(smem_gen + (wy0 + ((uint32_t)it.bseq % kWyTiles) * kWyTile - smem_base)) + half * 32;
      if (it.bseq >= 0) mbar_wait(smem_u32(&wy_full[(uint32_t)it.bseq % kWyTiles]), ((uint32_t)it.bseq / kWyTiles) & 1u, 11);
      float acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = 0.f;
      for (int y = it.ya; y < it.yb; ++y, ++rc) {
        if (ew == 0) { PROG(5, rc); PROG(6, iseq); }
        const uint32_t ab = rc & 1u;
        const float4* wrow = reinterpret_cast<const float4*>(wy_s + (y - it.ya) * kNR);
        mbar_wait(smem_u32(&acc_full[ab]), (rc >> 1) & 1u, 10);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + ab * kAccCols + (uint32_t)half * 32u, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&acc_empty[ab]));
        if (lv.mode == MODE_ROW) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 wv = wrow[k];
            acc[4 * k + 0] = fmaf(wv.x, __uint_as_float(v[4 * k + 0]), acc[4 * k + 0]);
            acc[4 * k + 1] = fmaf(wv.y, __uint_as_float(v[4 * k + 1]), acc[4 * k + 1]);
            acc[4 * k + 2] = fmaf(wv.z, __uint_as_float(v[4 * k + 2]), acc[4 * k + 2]);
            acc[4 * k + 3] = fmaf(wv.w, __uint_as_float(v[4 * k + 3]), acc[4 * k + 3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) acc[k] = __fadd_rn(acc[k], __uint_as_float(v[k]));
        }
      }
      if (ew == 0) TRACE(11);
      // ---- the item's [ROI x channel] partial ----
      const int nmine = min(max(it.nroi - half * 32, 0), 32);
      const int pair_items = lv.items;
      float* obase = p.out + (long long)it.l * kC + c;
      if (pair_items == 1) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < nmine) obase[(long long)__ldg(p.perm + it.start + half * 32 + k) * 4 * kC] = acc[k];
      } else {
        float* pp = p.partial + ((long long)it.slot * kNR + half * 32) * kC + c;
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < nmine) __stcg(pp + k * kC, acc[k]);
        __threadfence();
        if (ew == 0) TRACE(12);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) s_last = atomicAdd(&p.pair_ctr[it.g * 4 + it.l], 1);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ew == 0) TRACE(13);
        if (s_last == pair_items - 1) {                       // last item of the (group, level): add the partials in item order
          __threadfence();
          // thread = 4 channels (c4) x 8 ROI slots (r = rg + 8 k): 16-byte loads, two partials in flight, item order kept
          const int c4 = (et & 31) * 4, rg = et >> 5;
          const float4* p0 = reinterpret_cast<const float4*>(p.partial + (((long long)it.g * p.items_per_group + lv.item0) * kNR + rg) * kC + c4);
          for (int kh = 0; kh < 8; kh += 4) {                 // 4 ROI slots at a time: 12 float4 live
            float4 sum[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) sum[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < pair_items; i += 2) {
              float4 a[4], b[4];
              const bool two = i + 1 < pair_items;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const bool ok = rg + 8 * (kh + k) < it.nroi;
                a[k] = ok ? __ldcg(p0 + ((long long)i * kNR + 8 * (kh + k)) * (kC / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                b[k] = ok && two ? __ldcg(p0 + ((long long)(i + 1) * kNR + 8 * (kh + k)) * (kC / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                sum[k].x = __fadd_rn(sum[k].x, a[k].x); sum[k].y = __fadd_rn(sum[k].y, a[k].y);
                sum[k].z = __fadd_rn(sum[k].z, a[k].z); sum[k].w = __fadd_rn(sum[k].w, a[k].w);
                if (two) {
                  sum[k].x = __fadd_rn(sum[k].x, b[k].x); sum[k].y = __fadd_rn(sum[k].y, b[k].y);
                  sum[k].z = __fadd_rn(sum[k].z, b[k].z); sum[k].w = __fadd_rn(sum[k].w, b[k].w);
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (rg + 8 * (kh + k) < it.nroi)
                *reinterpret_cast<float4*>(p.out + (long long)__ldg(p.perm + it.start + rg + 8 * (kh + k)) * 4 * kC + it.l * kC + c4) = sum[k];
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");      // s_last is rewritten by the next item
      }
      __syncwarp();
      if (ew == 0) TRACE(14);
      if (lane == 0) mbar_arrive(smem_u32(&item_empty[islot]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// [N*C][H][W] fp32 (LIN: [N*C][1][H*W]), box = [128 channels][1 row][32 elements], 128-byte swizzle
bool make_level_map(CUtensorMap* m, const float* base, long long rows, int Hv, int Wv) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)Wv, (cuuint64_t)Hv, (cuuint64_t)rows};
  const cuuint64_t gstr[2] = {(cuuint64_t)Wv * 4ull, (cuuint64_t)Wv * (cuuint64_t)Hv * 4ull};
  const cuuint32_t box[3] = {(cuuint32_t)kKC, 1u, (cuuint32_t)kC};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Plan {
  TcPool kp;
  size_t off_hdr, off_grp, off_perm, off_pair, off_ranges, off_wtab, off_btab, off_partial, total;
  size_t weights_smem;
  int any_tc;
};

// Shapes only (never the ROIs): which levels run on the tensor cores, band sizes, workspace layout.
bool make_plan(const int Hl[4], const int Wl[4], int N, int C, int R, Plan& pl) {
  TcPool& kp = pl.kp;
  pl.any_tc = 0;
  if (C != kC || N < 1 || N > kMaxFrames || R < 1) return false;
  if ((long long)N * C > 0x7fffffffLL) return false;
  long long chunks_per_group = 0;
  int wy_off = 0;
  pl.weights_smem = 0;
  for (int l = 0; l < 4; ++l) {
    LevelInfo& lv = kp.lv[l];
    lv.H = Hl[l]; lv.W = Wl[l];
    lv.Wp = (Wl[l] + 3) / 4 * 4;
    lv.mode = MODE_NONE;
    const long long hw = (long long)Hl[l] * Wl[l];
    if ((size_t)kNR * (Hl[l] + Wl[l]) * 4 <= 160 * 1024) {
      if (Wl[l] % 4 == 0 && Wl[l] >= 4) lv.mode = MODE_ROW;
      else if (hw % 4 == 0 && hw <= kSegW) lv.mode = MODE_LIN;
    }
    lv.Hv = lv.mode == MODE_LIN ? 1 : Hl[l];
    lv.Wv = lv.mode == MODE_LIN ? (int)hw : Wl[l];
    lv.nseg = (lv.Wv + kSegW - 1) / kSegW;
    if (lv.mode != MODE_NONE) {
      pl.any_tc = 1;
      chunks_per_group += (long long)lv.Hv * ((lv.Wv + kKC - 1) / kKC);
      pl.weights_smem = std::max(pl.weights_smem, (size_t)kNR * (Hl[l] + Wl[l]) * 4);
    }
    lv.wy_off = wy_off; wy_off += Hl[l] * kNR;
  }
  if (!pl.any_tc) return false;
  int bt_off = 0;
  for (int l = 0; l < 4; ++l) {
    LevelInfo& lv = kp.lv[l];
    lv.nxc = lv.mode == MODE_NONE ? 0 : (lv.Wv + kKC - 1) / kKC;
    lv.bt_off = bt_off; bt_off += lv.nxc * 2 * kNR * kKC;
  }
  kp.btab_stride = bt_off;
  int wx_off = wy_off;
  for (int l = 0; l < 4; ++l) { kp.lv[l].wx_off = wx_off; wx_off += kNR * kp.lv[l].Wp; }
  kp.wtab_stride = (wx_off + 3) / 4 * 4;
  // ~4 work items per SM at one group per frame; items between 8 and 64 chunks
  int per_sm = 4;
  if (const char* e = getenv("DMM_K5_ITEMS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= 64) per_sm = v; }   // tuning knob
  long long target = chunks_per_group * N / (kNumSMs * (long long)per_sm);
  target = std::min<long long>(64, std::max<long long>(8, target));
  int item0 = 0;
  for (int l = 0; l < 4; ++l) {
    LevelInfo& lv = kp.lv[l];
    if (lv.mode == MODE_NONE) { lv.rows_per_band = 1; lv.nband = 0; lv.items = 0; lv.item0 = item0; continue; }
    const int seg_chunks = (std::min(lv.Wv, kSegW) + kKC - 1) / kKC;
    lv.rows_per_band = (int)std::min<long long>(kMaxBandRows, std::max<long long>(1, target / seg_chunks));
    lv.nband = (lv.Hv + lv.rows_per_band - 1) / lv.rows_per_band;
    lv.rows_per_band = (lv.Hv + lv.nband - 1) / lv.nband;      // equal bands
    lv.items = lv.nseg * lv.nband;
    lv.item0 = item0; item0 += lv.items;
  }
  kp.items_per_group = item0;
  kp.N = N; kp.R = R; kp.C = C; kp.tables_only = 0;
  kp.Gb = std::min(N, R) + R / kNR;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
  pl.off_hdr = take(64);
  pl.off_grp = take((size_t)kp.Gb * 3 * sizeof(int));
  pl.off_perm = take((size_t)R * sizeof(int));
  pl.off_pair = take((size_t)kp.Gb * 4 * sizeof(int));
  pl.off_ranges = take((size_t)kp.Gb * 16 * sizeof(int));
  pl.off_wtab = take((size_t)kp.Gb * kp.wtab_stride * sizeof(float));
  pl.off_btab = take((size_t)kp.Gb * kp.btab_stride * sizeof(float));
  pl.off_partial = take((size_t)kp.Gb * kp.items_per_group * kNR * kC * sizeof(float));
  pl.total = o;
  if (pl.total > ((size_t)2 << 30)) return false;            // absurd ROI counts: leave them to the gather kernel
  return true;
}


// ---------------------------------------------------------------------------------------------------------------------
// backward: gF[n, c, y, x] = sum over the frame's ROIs r of  gout[r, l*C + c] * wy[r, y] * wx[r, x]
// A GATHER over the ROIs of the frame in bucket order -- every gradient element is written exactly once, by one thread,
// with a fixed summation order: deterministic, no atomics, no zero-initialised output (the forward's SIMT sibling
// scattered with atomicAdd).  CTA = (frame, level, band of rows, chunk of <= 128 channels); per row the ROIs whose wy is
// non-zero are compacted and A[r][c] = gout[r][c] * wy[r][y] staged once; lanes run along x (coalesced stores), each
// thread carries 8 channels: 8 FMAs per (wx load + two broadcast LDS.128).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 512;
constexpr int kBwdCT = 8;                       // channels per thread

struct BwdParams {
  TcPool tp;                                    // groups, perm, weight tables, level shapes
  const float* gout;                            // [R][4*C]
  float* gfeat[4];
  int rows_per_band[4], nband[4], band0[4];     // CTA decomposition per level
};

__global__ void __launch_bounds__(kBwdThreads) roi_pool_bwd_kernel(const BwdParams bp) {
  extern __shared__ float bsm[];
  const TcPool& p = bp.tp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // decode (level, band) from blockIdx.x, frame from blockIdx.y, channel chunk from blockIdx.z
  int l = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) if ((int)blockIdx.x >= bp.band0[k]) l = k;
  const int band = blockIdx.x - bp.band0[l], n = blockIdx.y, c0 = blockIdx.z * kC;
  const LevelInfo& lv = p.lv[l];
  const int H = lv.H, W = lv.W, Wp = lv.Wp, C = p.C;
  const int cn = min(kC, C - c0);
  const int y0 = band * bp.rows_per_band[l], y1 = min(y0 + bp.rows_per_band[l], H);
  float* gsm = bsm;                              // gout tile   [64][128]
  float* asm_ = gsm + kNR * kC;                  // A           [64][128]  (compacted rows)
  float* wxs = asm_ + kNR * kC;                  // wx          [64][Wp]
  __shared__ int s_act[kNR], s_nact, s_g0, s_g1;
  if (tid == 0) {                                // groups of frame n: contiguous in the bucket order
    const int G = p.hdr[0];
    int a = 0, b = G;
    while (a < b) { const int m = (a + b) >> 1; if (p.grp_frame[m] < n) a = m + 1; else b = m; }
    int e = a;
    while (e < G && p.grp_frame[e] == n) ++e;
    s_g0 = a; s_g1 = e;
  }
  __syncthreads();
  const int g0 = s_g0, g1 = s_g1;
  float* out = bp.gfeat[l] + ((long long)n * C + c0) * H * W;
  const int nxc = (W + 31) / 32, ncg = (cn + kBwdCT - 1) / kBwdCT;   // tasks per row: x chunks x channel groups
  for (int y = y0; y < y1; ++y) {
    // accumulators live across the frame's groups: a task's 8 sums per thread; tasks are walked twice per group pass,
    // so they are kept in shared-memory-free registers only when there is one group; the general case re-adds from global
    for (int g = g0; g < max(g1, g0 + 1); ++g) {
      const bool have = g < g1;
      const int nroi = have ? p.grp_count[g] : 0, start = have ? p.grp_start[g] : 0;
      const float* tab = p.wtab + (long long)g * p.wtab_stride;
      __syncthreads();
      if (have && (y == y0 || g1 - g0 > 1)) {      // tiles of this group (once per band when the frame has a single group)
        for (int i = tid; i < kNR * kC; i += kBwdThreads) {
          const int r = i / kC, c = i % kC;
          gsm[i] = (r < nroi && c < cn) ? __ldg(bp.gout + (long long)__ldg(p.perm + start + r) * 4 * C + (long long)l * C + c0 + c) : 0.f;
        }
        for (int i = tid; i < kNR * Wp; i += kBwdThreads) wxs[i] = __ldg(tab + lv.wx_off + i);
      }
      __syncthreads();
      if (warp == 0) {                           // ROIs with weight on this row, compacted in ascending slot order (ballots)
        const bool f0 = lane < nroi && __ldg(tab + lv.wy_off + y * kNR + lane) != 0.f;
        const bool f1 = lane + 32 < nroi && __ldg(tab + lv.wy_off + y * kNR + lane + 32) != 0.f;
        const unsigned m0 = __ballot_sync(0xffffffffu, f0), m1 = __ballot_sync(0xffffffffu, f1), lt = (1u << lane) - 1u;
        if (f0) s_act[__popc(m0 & lt)] = lane;
        if (f1) s_act[__popc(m0) + __popc(m1 & lt)] = lane + 32;
        if (lane == 0) s_nact = __popc(m0) + __popc(m1);
      }
      __syncthreads();
      const int nact = s_nact;
      for (int i = tid; i < nact * kC; i += kBwdThreads) {
        const int a = i / kC, c = i % kC, r = s_act[a];
        asm_[a * kC + c] = __fmul_rn(gsm[r * kC + c], __ldg(tab + lv.wy_off + y * kNR + r));
      }
      __syncthreads();
      for (int task = warp; task < nxc * ncg; task += kBwdThreads / 32) {
        const int xc = task % nxc, cg = task / nxc;
        const int x = xc * 32 + lane;
        float acc[kBwdCT];
#pragma unroll
        for (int k = 0; k < kBwdCT; ++k) acc[k] = 0.f;
        if (x < W) {
          for (int a = 0; a < nact; ++a) {
            const float w = wxs[s_act[a] * Wp + x];
            const float4 a0 = *reinterpret_cast<const float4*>(asm_ + a * kC + cg * kBwdCT);
            const float4 a1 = *reinterpret_cast<const float4*>(asm_ + a * kC + cg * kBwdCT + 4);
            acc[0] = fmaf(a0.x, w, acc[0]); acc[1] = fmaf(a0.y, w, acc[1]); acc[2] = fmaf(a0.z, w, acc[2]); acc[3] = fmaf(a0.w, w, acc[3]);
            acc[4] = fmaf(a1.x, w, acc[4]); acc[5] = fmaf(a1.y, w, acc[5]); acc[6] = fmaf(a1.z, w, acc[6]); acc[7] = fmaf(a1.w, w, acc[7]);
          }
#pragma unroll
          for (int k = 0; k < kBwdCT; ++k) {
            const int c = cg * kBwdCT + k;
            if (c < cn) {
              float* dst = out + ((long long)c * H + y) * W + x;
              *dst = g == g0 ? acc[k] : __fadd_rn(*dst, acc[k]);   // frames with > 64 ROIs: groups added in bucket order
            }
          }
        }
      }
    }
  }
}

struct BwdPlan {
  BwdParams bp;
  size_t off_hdr, off_grp, off_perm, off_pair, off_wtab, total, weights_smem, smem;
  int bands_total;
};

bool make_bwd_plan(const int Hl[4], const int Wl[4], int N, int C, int R, BwdPlan& pl) {
  TcPool& kp = pl.bp.tp;
  if (N < 1 || N > kMaxFrames || N > 65535 || R < 1 || C < 1 || (C + kC - 1) / kC > 65535) return false;
  int wy_off = 0, wmax = 0;
  pl.weights_smem = 0;
  for (int l = 0; l < 4; ++l) {
    LevelInfo& lv = kp.lv[l];
    lv = LevelInfo();
    lv.H = Hl[l]; lv.W = Wl[l]; lv.Wp = (Wl[l] + 3) / 4 * 4;
    lv.mode = MODE_NONE;
    if ((size_t)kNR * (Hl[l] + Wl[l]) * 4 > 160 * 1024 || lv.Wp > 512) return false;
    pl.weights_smem = std::max(pl.weights_smem, (size_t)kNR * (Hl[l] + Wl[l]) * 4);
    wmax = std::max(wmax, lv.Wp);
    lv.wy_off = wy_off; wy_off += Hl[l] * kNR;
  }
  int wx_off = wy_off;
  for (int l = 0; l < 4; ++l) { kp.lv[l].wx_off = wx_off; wx_off += kNR * kp.lv[l].Wp; }
  kp.wtab_stride = (wx_off + 3) / 4 * 4;
  kp.btab_stride = 0; kp.items_per_group = 0;
  kp.N = N; kp.R = R; kp.C = C; kp.tables_only = 1;
  kp.Gb = std::min(N, R) + R / kNR;
  pl.smem = ((size_t)2 * kNR * kC + (size_t)kNR * wmax) * sizeof(float);
  // bands: ~8 CTAs per SM over all frames; at least 2 rows per band
  long long rows_total = 0;
  for (int l = 0; l < 4; ++l) rows_total += Hl[l];
  const long long cchunks = (C + kC - 1) / kC;
  long long rpb = std::max<long long>(2, rows_total * N * cchunks / (kNumSMs * 8));
  int b0 = 0;
  for (int l = 0; l < 4; ++l) {
    pl.bp.rows_per_band[l] = (int)std::min<long long>(rpb, Hl[l]);
    pl.bp.nband[l] = (Hl[l] + pl.bp.rows_per_band[l] - 1) / pl.bp.rows_per_band[l];
    pl.bp.band0[l] = b0; b0 += pl.bp.nband[l];
  }
  pl.bands_total = b0;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
  pl.off_hdr = take(64);
  pl.off_grp = take((size_t)kp.Gb * 3 * sizeof(int));
  pl.off_perm = take((size_t)R * sizeof(int));
  pl.off_pair = take((size_t)kp.Gb * 4 * sizeof(int));
  pl.off_wtab = take((size_t)kp.Gb * kp.wtab_stride * sizeof(float));
  pl.total = o;
  return pl.total <= ((size_t)2 << 30);
}

}  // namespace

size_t roi_pool_bwd_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R) {
  BwdPlan pl;
  return make_bwd_plan(Hl, Wl, N, C, R, pl) ? pl.total : 0;
}

// Deterministic backward.  Returns DMM_OK when launched (g_feat is then fully overwritten), -1 when the shapes are outside
// its envelope (the caller falls back to the atomic scatter kernel, which needs zero-initialised g_feat), or DMM_ERR_*.
int roi_pool_bwd_try_launch(const float* g_out, const int Hl[4], const int Wl[4], int N, int C, const float* rois, int R,
                            float* const g_feat[4], void* workspace, size_t workspace_bytes, cudaStream_t st) {
  BwdPlan pl;
  if (!workspace || ((uintptr_t)workspace & 255u) || !make_bwd_plan(Hl, Wl, N, C, R, pl) || workspace_bytes < pl.total) return -1;
  TcPool& kp = pl.bp.tp;
  uint8_t* ws = (uint8_t*)workspace;
  kp.rois = rois; kp.out = nullptr;
  kp.hdr = (int*)(ws + pl.off_hdr);
  kp.grp_frame = (int*)(ws + pl.off_grp);
  kp.grp_start = kp.grp_frame + kp.Gb;
  kp.grp_count = kp.grp_start + kp.Gb;
  kp.perm = (int*)(ws + pl.off_perm);
  kp.pair_ctr = (int*)(ws + pl.off_pair);
  kp.ranges = nullptr; kp.btab = nullptr; kp.partial = nullptr; kp.trace = nullptr;
  kp.wtab = (float*)(ws + pl.off_wtab);
  pl.bp.gout = g_out;
  for (int l = 0; l < 4; ++l) pl.bp.gfeat[l] = g_feat[l];
  static size_t attr_w[64] = {0}, attr_b[64] = {0};
  int dev = 0;
  DMM_CUDA_TRY(cudaGetDevice(&dev));
  const int di = dev & 63;
  tc_group_kernel<<<1, 1024, 0, st>>>(kp);
  if (pl.weights_smem > 48 * 1024 && attr_w[di] < pl.weights_smem) {
    DMM_CUDA_TRY(cudaFuncSetAttribute(tc_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.weights_smem));
    attr_w[di] = pl.weights_smem;
  }
  tc_weights_kernel<<<dim3(kp.Gb, 4), 256, pl.weights_smem, st>>>(kp);
  if (pl.smem > 48 * 1024 && attr_b[di] < pl.smem) {
    DMM_CUDA_TRY(cudaFuncSetAttribute(roi_pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    attr_b[di] = pl.smem;
  }
  roi_pool_bwd_kernel<<<dim3(pl.bands_total, N, (C + kC - 1) / kC), kBwdThreads, pl.smem, st>>>(pl.bp);
  return check_launch();
}

namespace {
}  // namespace

size_t roi_pool_tc_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R) {
  Plan pl;
  return make_plan(Hl, Wl, N, C, R, pl) ? pl.total : 0;
}

// Launches the tensor-core path for the levels it can take.  Returns DMM_OK and sets *level_mask (bit l = level l done
// here), -1 when nothing can run here (caller uses the SIMT kernel for everything), or a DMM_ERR_* code.
int roi_pool_tc_try_launch(const float* const feat[4], const int Hl[4], const int Wl[4], int N, int C, const float* rois, int R,
                           float* out, void* workspace, size_t workspace_bytes, int* level_mask, cudaStream_t st) {
  *level_mask = 0;
  Plan pl;
  if (!workspace || !make_plan(Hl, Wl, N, C, R, pl) || workspace_bytes < pl.total) return -1;
  if ((uintptr_t)workspace & 255u) return -1;
  TcPool& kp = pl.kp;
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  int mask = 0;
  for (int l = 0; l < 4; ++l) {
    LevelInfo& lv = kp.lv[l];
    if (lv.mode == MODE_NONE) continue;
    if (((uintptr_t)feat[l] & 15u) || !make_level_map(&maps[l], feat[l], (long long)N * C, lv.Hv, lv.Wv)) return -1;
    mask |= 1 << l;
  }
  if (!mask) return -1;
  *level_mask = mask;
  uint8_t* ws = (uint8_t*)workspace;
  kp.rois = rois; kp.out = out;
  kp.hdr = (int*)(ws + pl.off_hdr);
  kp.grp_frame = (int*)(ws + pl.off_grp);
  kp.grp_start = kp.grp_frame + kp.Gb;
  kp.grp_count = kp.grp_start + kp.Gb;
  kp.perm = (int*)(ws + pl.off_perm);
  kp.pair_ctr = (int*)(ws + pl.off_pair);
  kp.ranges = (int*)(ws + pl.off_ranges);
  kp.wtab = (float*)(ws + pl.off_wtab);
  kp.btab = (float*)(ws + pl.off_btab);
  kp.partial = (float*)(ws + pl.off_partial);
  kp.trace = nullptr;
#ifdef DMM_TC_DEBUG
  static long long* trace_buf = nullptr;
  if (getenv("DMM_K5_TRACE_FILE")) {
    if (!trace_buf) cudaMalloc(&trace_buf, 4096 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 4096 * sizeof(long long), st);
    kp.trace = trace_buf;
  }
#endif
  tc_group_kernel<<<1, 1024, 0, st>>>(kp);
  // function attributes are per device and sticky: set them once per (process, device), not on every call
  static size_t attr_weights[64] = {0};
  static bool attr_main[64] = {false};
  int dev = 0;
  DMM_CUDA_TRY(cudaGetDevice(&dev));
  const int di = dev & 63;
  if (pl.weights_smem > 48 * 1024 && attr_weights[di] < pl.weights_smem) {
    DMM_CUDA_TRY(cudaFuncSetAttribute(tc_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.weights_smem));
    attr_weights[di] = pl.weights_smem;
  }
  tc_weights_kernel<<<dim3(kp.Gb, 4), 256, pl.weights_smem, st>>>(kp);
  if (!attr_main[di]) {
    DMM_CUDA_TRY(cudaFuncSetAttribute(roi_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem));
    attr_main[di] = true;
  }
  roi_pool_tc_kernel<<<kNumSMs, kThreads, kDynSmem, st>>>(kp, maps[0], maps[1], maps[2], maps[3]);
#ifdef DMM_TC_DEBUG
  if (kp.trace) {
    static long long host[4096];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, kp.trace, sizeof(host), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(getenv("DMM_K5_TRACE_FILE"), "w")) {
      for (int i = 0; i < 2048 && host[2 * i]; ++i) fprintf(f, "%lld %lld\n", host[2 * i], host[2 * i + 1]);
      fclose(f);
    }
  }
#endif
  return check_launch();
}

}  // namespace dmm
