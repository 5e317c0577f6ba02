/*
 * dmm_b200.h -- C ABI of libdmm_b200.so: the sm_100a kernels behind DMM-Net's matching layer.
 *
 * Drop-in boundary.  The reference has no native code; what a maintainer binds
 * instead of the torch op sequences below is this library (ctypes stub: INTEGRATION.md).
 * Every entry point
 *   - takes raw DEVICE pointers, explicit sizes/strides, scalars and a CUDA stream (cudaStream_t cast to void*);
 *   - allocates nothing the caller can see (workspace is passed in, size from the *_workspace_bytes query);
 *   - is asynchronous on `stream`, re-entrant, keeps no global mutable state;
 *   - returns 0 on success or a DMM_ERR_* code (no exceptions).  Python keeps the reference's AssertionError
 *     shape checks (dmm/utils/checker.py) above this ABI.
 *
 * Shapes use the reference's names: P proposals, O templates, H*W = HW pixels, D feature channels,
 * B independent (video, frame) problems per launch.  All tensors are dense row-major fp32 unless noted.
 * Optional per-problem counts n_prop[B] / n_tmpl[B] (device int32, may be NULL) say how many of the P / O
 * rows of problem b are real; the rest is padding (dmm_model.py:117-125 selects the first O valid templates).
 * MS = max(P, O+1) is the column stride of every [O x m] solver matrix: problems with P <= O are padded with
 * zero-similarity dummy proposals to O+1 columns (match_model.py:109-113).
 */
#ifndef DMM_B200_H_
#define DMM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMM_B200_VERSION 0x000200 /* 0.2.0 */

enum {
  DMM_OK = 0,
  DMM_ERR_INVALID_ARGUMENT = 1, /* NULL where data is required, negative size, misuse */
  DMM_ERR_UNSUPPORTED_SHAPE = 2, /* beyond the compiled limits (see dmm_b200_limits) */
  DMM_ERR_WORKSPACE_TOO_SMALL = 3,
  DMM_ERR_CUDA = 4 /* a CUDA runtime call failed; dmm_b200_last_cuda_error() has the cudaError_t */
};

/* ---- library queries -------------------------------------------------------------------------------- */
int dmm_b200_version(void);                 /* DMM_B200_VERSION */
const char* dmm_b200_arch(void);            /* "sm_100a" */
const char* dmm_b200_error_string(int code);
int dmm_b200_last_cuda_error(void);         /* thread-local cudaError_t of the last DMM_ERR_CUDA */
/* out[0]=max O per solver problem, out[1]=max padded columns m per solver problem, out[2]=max D, out[3]=SM count used for grid sizing (0 if no device) */
int dmm_b200_limits(int* out4);

/* ---- K1: pairwise binary-mask IoU ---------------------------------------------------------------------
 * Replaces compute_iou_binary_mask_2D on the expanded [O*P, HW] pairs (match_helper.py:9-28 called from
 * match_model.py:83-89) and, with tmpl2 = targets, the second IoU build of compute_matching_loss
 * (match_helper.py:30-42) in the same pass over the proposals.
 *   iou[b,o,p] = |A_o & B_p| / (float(|A_o | B_p|) + 1e-6f),   bits = (mask > 0.5f)      -- bit-exact
 * prop [B][P][HW] (batch stride prop_bstride elements), tmpl [B][O][HW], tmpl2 optional (NULL) [B][O][HW].
 * Outputs (each may be NULL): iou/iou2 [B][O][P]; sim [B][O][P] = cos*w_cos + iou*w_iou (match_model.py:90,
 * separate fp32 roundings, no FMA) when cos != NULL; counts [B][3 slots] see below.
 * counts (optional, int32): [B][O*P + O + P] = inter[o][p], area_tmpl[o], area_prop[p] for the FIRST template set.
 */
size_t dmm_mask_iou_workspace_bytes(int B, int P, int O, int HW, int two_template_sets);
int dmm_mask_iou_pairwise(const float* prop, long long prop_bstride, const float* tmpl, long long tmpl_bstride,
                          const float* tmpl2, long long tmpl2_bstride, int B, int P, int O, int HW,
                          const int* n_prop, const int* n_tmpl, float* iou, float* iou2, const float* cos,
                          float w_cos, float w_iou, float* sim, int* counts, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Same, with the proposals of problem b at prop_ptrs[b] (device array of B device pointers, each [n_prop[b] or P][HW]):
 * the per-video proposal tensors of dmm_model.py:111 are used in place, without the torch.stack copy (23 MB per video
 * and frame at the headline size).  ptrs_aligned16 != 0 promises that every pointer is 16-byte aligned. */
int dmm_mask_iou_pairwise_ptrs(const float* const* prop_ptrs, int ptrs_aligned16, const float* tmpl,
                               long long tmpl_bstride, const float* tmpl2, long long tmpl2_bstride, int B, int P, int O,
                               int HW, const int* n_prop, const int* n_tmpl, float* iou, float* iou2, const float* cos,
                               float w_cos, float w_iou, float* sim, int* counts, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ---- bit-packed masks ----------------------------------------------------------------
 * bit i of word j of a row = (pixel 32*j + i) > 0.5f, zero past the row end; a row is dmm_packed_words(HW) uint32.
 * The IoU needs only these bits, so a producer that packs once moves 32x fewer bytes through K1:
 *   dmm_mask_pack_bits      device fp32 [rows][HW] -> device bits [rows][words]
 *   dmm_host_pack_masks     HOST fp32 -> HOST bits, multi-threaded with plain std::threads (threads <= 0: all cores), so that masks held in
 *                           host memory cross PCIe as 0.86 MB instead of 27.5 MB per match
 *   dmm_mask_iou_pairwise_packed   K1 on packed rows: same outputs, bit-identical to dmm_mask_iou_pairwise. */
long long dmm_packed_words(long long HW);
int dmm_host_pack_masks(const float* src_host, long long rows, long long HW, uint32_t* dst_host, int threads);
/* two row sets of the same HW (a problem's proposal and template masks) packed by ONE thread team */
int dmm_host_pack_masks2(const float* src_a, long long rows_a, uint32_t* dst_a, const float* src_b, long long rows_b,
                         uint32_t* dst_b, long long HW, int threads);
/* Streaming-read bandwidth (GB/s, best of `reps` passes) of a host buffer with the packer's parked thread team: the
 * host-DRAM roofline of the host-buffer entry (bench.py e2e.host_dram_read_gbs). */
int dmm_host_read_bandwidth(const void* src, long long bytes, int threads, int reps, double* gbs);
int dmm_mask_pack_bits(const float* masks, long long rows, int HW, uint32_t* bits, void* stream);
size_t dmm_mask_iou_packed_workspace_bytes(int B, int P, int O, int words, int two_template_sets);
int dmm_mask_iou_pairwise_packed(const uint32_t* prop_bits, long long prop_bstride_words, const uint32_t* tmpl_bits,
                                 long long tmpl_bstride_words, const uint32_t* tmpl2_bits,
                                 long long tmpl2_bstride_words, int B, int P, int O, int words, const int* n_prop,
                                 const int* n_tmpl, float* iou, float* iou2, const float* cos, float w_cos,
                                 float w_iou, float* sim, int* counts, void* workspace, size_t workspace_bytes,
                                 void* stream);

/* Row-paired IoU: a[N][M] vs b[N][M] -> iou[N]  (compute_iou_binary_mask_2D itself; callers trainer.py:190,298). */
size_t dmm_mask_iou_rowwise_workspace_bytes(int N, int M);
int dmm_mask_iou_rowwise(const float* a, const float* b, int N, int M, float* iou, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- K2: pairwise cosine ------------------------------------------------------------------------------
 * Replaces get_cosine_score (match_helper.py:51-64) averaged over the T template-feature sets
 * (match_model.py:72-76).  tmpl_feat [B][T][O][D], prop_feat [B][P][D] -> cos [B][O][P].
 * cos_t[o,p] = <q,k> / (max(|q|,eps) * max(|k|,eps)), eps = 1e-8 (F.cosine_similarity).
 */
int dmm_cosine_pairwise(const float* tmpl_feat, const float* prop_feat, int B, int T, int P, int O, int D,
                        const int* n_prop, const int* n_tmpl, float eps, float* cos, void* stream);
/* Same with an explicit kernel choice.  DMM_COSINE_TC: the contraction runs on the tensor cores (tcgen05.mma kind::tf32
 * as a 3xTF32 split with fp32 accumulation in TMEM, features streamed by TMA; |err| ~3e-6 of an fp64 reference);
 * envelope T == 1, P <= 64, O <= 16, D >= 32, D % 4 == 0, 16-byte aligned features, else DMM_ERR_UNSUPPORTED_SHAPE.
 * DMM_COSINE_SIMT: fp32 FFMA kernel, any shape (|err| ~3e-7).  DMM_COSINE_AUTO (what dmm_cosine_pairwise does): TC
 * inside the envelope, SIMT outside; the environment variable DMM_K2_IMPL=simt forces SIMT for dmm_cosine_pairwise. */
enum { DMM_COSINE_AUTO = 0, DMM_COSINE_SIMT = 1, DMM_COSINE_TC = 2 };
int dmm_cosine_pairwise_impl(const float* tmpl_feat, const float* prop_feat, int B, int T, int P, int O, int D,
                             const int* n_prop, const int* n_tmpl, float eps, float* cos, int impl, void* stream);
/* d(loss)/d(features) from g_cos [B][O][P]; g_tmpl_feat [B][T][O][D], g_prop_feat [B][P][D] are overwritten.
 * cos_fwd (optional, may be NULL): the forward's output; used instead of recomputing the cosines when T == 1. */
int dmm_cosine_pairwise_bwd(const float* g_cos, const float* cos_fwd, const float* tmpl_feat, const float* prop_feat,
                            int B, int T, int P, int O, int D, const int* n_prop, const int* n_tmpl, float eps,
                            float* g_tmpl_feat, float* g_prop_feat, void* stream);

/* ---- K3: relaxed matching solver + assignment head ----------------------------------------------------
 * Replaces relax_matching (relax_match.py:36-105: greedy init, gradient step, Dykstra sweeps over
 * {X>=0} n {col sums<=1} n {row sums==1}, both exact-equality early exits) and the [O x m] part of
 * match_with_first_frame (match_model.py:109-130,146-147): zero-column padding when P<=O, C=-sim,
 * R = mean of the PRE-projection iterates, logic = (R==rowmax) [is_test] or (R>0.01), Bmat = R*logic,
 * match_score = max_p(clamp(R,0,1)*sim), det_score = sum_p score_p*Bmat.
 * One CTA of 4 warps per problem; X, the three Dykstra increments, C and the running sum of iterates stay in registers.
 *
 * in : mat [B][O][P]  (a similarity if negate!=0, else used as the cost C directly), prop_score [B][P] or NULL
 * out: (any may be NULL)  R, X_final, Bmat, logic [B][O][MS];  match_score, det_score [B][O];
 *      n_list [B] = len(X_list) = 1 + outer steps run;
 *      xlist [B][max_iter+1][O][MS] every recorded iterate, cost [B][max_iter+1] (cost[0]=0) -- the
 *      free-function API returns them (relax_match.py:105);
 *      saved: opaque forward state for dmm_relax_solve_bwd, dmm_relax_saved_bytes(B,max_iter,proj_iter) bytes.
 * pad_to_square_plus_one: apply the P<=O padding rule (the layer does; the free function does not).
 */
size_t dmm_relax_saved_bytes(int B, int max_iter, int proj_iter);
int dmm_relax_solve(const float* mat, const float* prop_score, int B, int P, int O, const int* n_prop,
                    const int* n_tmpl, int max_iter, int proj_iter, float lr, int negate, int pad_rule,
                    int is_test, float* R, float* X_final, float* Bmat, float* logic, float* match_score,
                    float* det_score, int* n_list, float* xlist, float* cost, void* saved, void* stream);
/* Backward of the above w.r.t. `mat` and prop_score.  Cotangents (any may be NULL): g_R, g_Xfinal, g_Bmat [B][O][MS],
 * g_match_score, g_det_score [B][O].  Needs the forward's R, logic, mat, prop_score, n_list and `saved`.
 * Outputs: g_mat [B][O][P] (overwritten), g_prop_score [B][P] (overwritten, may be NULL). */
int dmm_relax_solve_bwd(const float* g_R, const float* g_Xfinal, const float* g_Bmat, const float* g_match_score,
                        const float* g_det_score, const float* mat, const float* prop_score, const float* R,
                        const float* logic, const int* n_list, const void* saved, int B, int P, int O,
                        const int* n_prop, const int* n_tmpl, int max_iter, int proj_iter, float lr, int negate,
                        int pad_rule, float* g_mat, float* g_prop_score, void* stream);

/* ---- K4: assignment apply -----------------------------------------------------------------------------
 * Replaces full_outmask = torch.mm(binary_Ridx_matched, proposed_mask2d) (match_model.py:144) and, with
 * row_map, the valid-row scatter torch.mm(FO_matrix, .) of dmm_model.py:133-135.
 * Bmat [B][O][MS] (columns >= P are padding and multiply zero rows), prop [B][P][HW] -> out [B][O_out][HW];
 * row o of problem b goes to output row row_map[b*O+o] (identity when row_map==NULL); when zero_fill!=0 every
 * output row not produced is written with zeros.  Only the non-zero coefficients are streamed.
 */
int dmm_assign_apply(const float* Bmat, const float* prop, long long prop_bstride, int B, int P, int O, int MS,
                     int HW, const int* n_prop, const int* n_tmpl, const int* row_map, int O_out, int zero_fill,
                     float* out, long long out_bstride, void* stream);
int dmm_assign_apply_ptrs(const float* Bmat, const float* const* prop_ptrs, int ptrs_aligned16, int B, int P, int O,
                          int MS, int HW, const int* n_prop, const int* n_tmpl, const int* row_map, int O_out,
                          int zero_fill, float* out, long long out_bstride, void* stream);
/* g_Bmat[b,o,p] = <g_out[b,row(o),:], prop[b,p,:]> for the entries selected by `logic` (others 0);
 * optional g_prop [B][P][HW] = Bmat^T g_out (overwritten) when the proposal masks need a gradient. */
size_t dmm_assign_apply_bwd_workspace_bytes(int B, int P, int O, int HW);
int dmm_assign_apply_bwd(const float* g_out, long long gout_bstride, const float* prop, long long prop_bstride,
                         const float* Bmat, const float* logic, int B, int P, int O, int MS, int HW,
                         const int* n_prop, const int* n_tmpl, const int* row_map, float* g_Bmat, float* g_prop,
                         void* workspace, size_t workspace_bytes, void* stream);

int dmm_assign_apply_bwd_ptrs(const float* g_out, long long gout_bstride, const float* const* prop_ptrs,
                              int ptrs_aligned16, const float* Bmat, const float* logic, int B, int P, int O, int MS,
                              int HW, const int* n_prop, const int* n_tmpl, const int* row_map, float* g_Bmat,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- K5: ROI mean pooling -----------------------------------------------------------------------------
 * Replaces FeatureExtractor.forward (feature_extractor.py:20-52): legacy ROIAlign(14x14, sampling_ratio 2,
 * scales 1/4..1/32) on each of 4 levels followed by the spatial mean, i.e. the separable contraction
 *   out[r, l*C + c] = sum_y sum_x wy[r,l,y] * wx[r,l,x] * F_l[b_r, c, y, x].
 * feat[l] is [N][C][Hl][Wl]; rois [R][5] = (batch index, x1, y1, x2, y2) in image pixels; out [R][4*C].
 *
 * Two implementations behind one entry: the tensor-core path (csrc/roi_pool_tc.cu: per frame, the ROIs are the N
 * dimension of a tcgen05 GEMM over TMA-streamed feature rows, so every feature byte is read once per frame) for C == 128
 * and every level with Wl % 4 == 0 (or Hl*Wl <= 128), and the SIMT gather kernel for anything else.  `workspace`
 * (dmm_roi_mean_pool_workspace_bytes, 256-byte aligned; may be NULL -> SIMT only) holds the ROI buckets, weight tables
 * and band partials.  impl: 0 auto, 1 SIMT only, 2 tensor-core path required (DMM_ERR_UNSUPPORTED_SHAPE otherwise).
 */
size_t dmm_roi_mean_pool_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R);
int dmm_roi_mean_pool(const float* const feat[4], const int Hl[4], const int Wl[4], int N, int C,
                      const float* rois, int R, float* out, void* workspace, size_t workspace_bytes, int impl,
                      void* stream);
/* Backward.  Two implementations: a deterministic GATHER (per frame, level, row: sum over the frame's ROIs in bucket
 * order; every gradient element written once, no atomics; needs `workspace`, dmm_roi_mean_pool_bwd_workspace_bytes) that
 * OVERWRITES g_feat, and the atomic scatter (any shape) that ACCUMULATES into a zero-initialised g_feat.
 * impl: 0 auto, 1 scatter only, 2 gather required.  *wrote_all (may be NULL) = 1 when g_feat was fully overwritten. */
size_t dmm_roi_mean_pool_bwd_workspace_bytes(const int Hl[4], const int Wl[4], int N, int C, int R);
int dmm_roi_mean_pool_bwd(const float* g_out, const int Hl[4], const int Wl[4], int N, int C, const float* rois,
                          int R, float* const g_feat[4], void* workspace, size_t workspace_bytes, int impl,
                          int* wrote_all, void* stream);

/* ---- K6: decoder mask-input pyramid -------------------------------------------------
 * Replaces, for every object at once, trainer.py:256-263 / evaluator.py:187-194:
 *   prev_m_inst = cat(prev_mask[:,t], ref_mask[:,t], init_pred_inst[:,t]) -> [B,3,H,W];
 *   nn.MaxPool2d((2,2), ceil_mode=True) applied 1+L times, the last L results kept.
 * prev / ref / init: [B][O][H][W] (batch strides in elements, object stride H*W).
 * out_levels: HOST array of L device pointers; level k (window 4<<k) is [O][B][3][hk][wk] with
 * hk = ceil(H / (4<<k)), wk = ceil(W / (4<<k)) (dmm_mask_pyramid_level_size), so out_levels[k] + t*B*3*hk*wk is the
 * reference's mask_lstm entry of object t (the reference list is reversed: coarsest first).  L <= 5.  Bit-exact.
 */
int dmm_mask_pyramid_level_size(int H, int W, int level, int* h_out, int* w_out);
int dmm_mask_pyramid(const float* prev, long long prev_bstride, const float* ref, long long ref_bstride,
                     const float* init, long long init_bstride, int B, int O, int H, int W, int L,
                     float* const* out_levels, void* stream);
/* Backward: g_out_levels[k] (may be NULL = zero cotangent) in the layout of out_levels; g_prev / g_ref / g_init
 * dense [B][O][H][W], overwritten, each may be NULL.  Gradients go to the first maximum of every 2x2 stage in
 * row-major order (max_pool2d_with_indices); deterministic, bit-equal to autograd through the chained pools. */
int dmm_mask_pyramid_bwd(const float* const* g_out_levels, const float* prev, long long prev_bstride, const float* ref,
                         long long ref_bstride, const float* init, long long init_bstride, int B, int O, int H, int W,
                         int L, float* g_prev, float* g_ref, float* g_init, void* stream);

/* ---- K7: merged label map ------------------------------------------------------------------------------------
 * Replaces evaluator.py:139-145: label[b,px] = argmax([1 - max_o m[b,o,px], m[b,0,px], ..., m[b,n-1,px]]) with
 * n = n_valid[b] (NULL: O) valid objects; first maximum wins; n == 0 gives 0.  masks [B][O][HW] -> label uint8 [B][HW].
 */
int dmm_merge_labels(const float* masks, long long bstride, int B, int O, int HW, const int* n_valid,
                     unsigned char* label, void* stream);

/* ---- K8: proposal mask paste ----------------------------------------------------------
 * Replaces Masker.forward_single_image / paste_mask_in_image / binmask_to_box (dmm/utils/masker.py:91-206) for all N
 * proposals of a batch of frames in one launch.  masks [N][M][M] soft, boxes [N][4] xyxy (image pixels), padding >= 1.
 * Outputs (each may be NULL):
 *   pasted [N][im_h][im_w] fp32 -- the zero-padded mask bilinearly resized (align_corners=False) to the expanded int
 *          box and pasted into a zero image; these are the dense proposal rows K1 / K4 read;
 *   bits   [N][dmm_packed_words(im_h*im_w)] -- (pasted > 0.5f) in the packed-row format of dmm_mask_iou_pairwise_packed;
 *   tight  [N][4] int64 -- binmask_to_box(pasted > thresh): xmin, ymin, xmax, ymax, or 0, 0, im_h, im_w when empty.
 * workspace: dmm_paste_masks_workspace_bytes(N) bytes, needed when tight != NULL.  M + 2*padding <= 64.
 * A box entirely outside the image pastes nothing (the reference raises there).
 */
size_t dmm_paste_masks_workspace_bytes(int N);
int dmm_paste_masks(const float* masks, const float* boxes, int N, int M, int padding, int im_h, int im_w, float thresh,
                    float* pasted, uint32_t* bits, long long* tight, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---- K10: fused paste + assignment apply ("lazy paste") -----------------------------------------------------------
 * out[b, row(o), :] = sum_p Bmat[b,o,p] * paste(masks[src_index[b,p]], boxes[src_index[b,p]])
 * = dmm_assign_apply on top of dmm_paste_masks without ever materialising the P pasted proposal masks: with K8's bit rows
 * feeding the packed K1 entry, the eval pipeline writes O*HW*4 bytes per frame instead of writing P*HW*4, reading them for
 * the IoU and reading the selected ones again for the apply.  Bit-identical to dmm_paste_masks followed by dmm_assign_apply.
 * Bmat [B][O][MS]; masks [Nsrc][M][M], boxes [Nsrc][4]; src_index [B][P] = row of masks/boxes behind column p of problem
 * b (< 0: none; this is where the NMS keep list plugs in); n_prop / n_tmpl / row_map / O_out / zero_fill as in
 * dmm_assign_apply; out [B][O_out][im_h][im_w].  Inference only (no backward).  P <= 128, B*O_out <= 65535.
 */
int dmm_paste_apply(const float* Bmat, const float* masks, const float* boxes, const int* src_index, int B, int P, int O,
                    int MS, int M, int padding, int im_h, int im_w, const int* n_prop, const int* n_tmpl,
                    const int* row_map, int O_out, int zero_fill, float* out, long long out_bstride, void* stream);
/* Backward of dmm_paste_apply w.r.t. the assignment: g_Bmat[b,o,p] = <g_out[b,row(o)], paste(detection src_index[b,p])> for
 * the entries with sel[b,o,p] != 0 (the solver's selection mask), 0 elsewhere.  The mask-head outputs carry no gradient
 * (offline proposals, dmm/modules/model_encoder.py).  Deterministic (fixed-order block reduction). */
int dmm_paste_apply_bwd(const float* g_out, long long gout_bstride, const float* sel, const float* masks,
                        const float* boxes, const int* src_index, int B, int P, int O, int MS, int M, int padding,
                        int im_h, int im_w, const int* n_prop, const int* n_tmpl, const int* row_map, int O_out,
                        float* g_Bmat, void* stream);

/* ---- K9: box NMS ---------------------------------------------------------------------------------------------
 * Replaces filter_results (dmm/utils/boxlist_ops.py:15-29) -> maskrcnn_benchmark.layers.nms (un-vendored): greedy,
 * score-descending (ties: lower index first), IoU with the legacy +1 pixel widths, suppress when IoU > thresh, then
 * keep at most max_keep (<= 0: all).  One CTA per frame.
 * boxes [F][n_max][4], scores [F][n_max], n_boxes [F] (NULL: n_max) -> keep [F][n_max] int64 (kept indices in score
 * order, -1 padded), n_keep [F].  n_max <= 1024.
 */
int dmm_box_nms(const float* boxes, const float* scores, const int* n_boxes, int F, int n_max, float thresh,
                int max_keep, long long* keep, int* n_keep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMM_B200_H_ */
