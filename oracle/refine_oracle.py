"""CPU oracle for the rows either side of the matching layer.  TEST INFRASTRUCTURE ONLY.

Same rules as ``oracle/match_oracle.py``: torch-CPU fp32 restatement of the reference op sequences, imported only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs; never by the product path.

Parity status
  mask_pyramid, merged_labels   PINNED: ``oracle/make_golden_refine.py`` executes the reference's own source lines
                                (dmm/modules/trainer.py:256-263, dmm/modules/evaluator.py:139-145, read from
                                /root/reference at generation time) on seeded inputs and stores what they produced
                                under ``tests/golden/refine_*.npz``; ``tests/test_refine_rows.py`` holds this file to them.
  paste_masks                   PINNED: the reference's ``dmm.utils.masker.paste_mask_in_image`` is imported (with the
                                un-vendored ``maskrcnn_benchmark`` names it pulls in stubbed: ``interpolate`` is
                                ``torch.nn.functional.interpolate``, which is what that wrapper forwards to) and its
                                outputs stored in ``tests/golden/paste_*.npz`` (``tests/test_paste_nms.py``: reproduced bit for bit).
  box_nms                       parity UNPINNED: the arithmetic lives in the un-vendored ``maskrcnn_benchmark.layers.nms``
                                (fork without a pinned commit, INSTALL.md:19-39); restated from its published algorithm
                                (greedy, score-descending, IoU with the legacy +1 pixel widths, suppress when IoU > thresh).

Reference lines restated:
  modules/trainer.py:256-263, modules/evaluator.py:187-194 -> mask_pyramid
  modules/evaluator.py:139-145                             -> merged_labels
  modules/trainer.py:189-196                               -> hard_iou_mean
  utils/masker.py:91-173                                   -> paste_masks
  utils/boxlist_ops.py:15-29 (+ maskrcnn_benchmark nms)    -> box_nms, filter_results
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F

from .match_oracle import rowwise_binary_iou


def mask_pyramid(prev_mask: torch.Tensor, ref_mask: torch.Tensor, init_pred: torch.Tensor, n_levels: int) -> List[torch.Tensor]:
    """[B,O,H,W] x3 -> n_levels tensors [O,B,3,hk,wk], finest (window 4) first.
    trainer.py:256-263: per object, cat the three masks as channels, MaxPool2d((2,2), ceil_mode=True) once, then once
    more per level, keeping each result."""
    B, O, H, W = init_pred.shape
    pool = torch.nn.MaxPool2d((2, 2), ceil_mode=True)
    per_level: List[List[torch.Tensor]] = [[] for _ in range(n_levels)]
    for t in range(O):
        m = torch.cat([prev_mask[:, t].reshape(B, 1, H * W), ref_mask[:, t].reshape(B, 1, H * W),
                       init_pred[:, t].reshape(B, 1, H * W)], dim=2).view(B, 3, H, W)
        m = pool(m)
        for k in range(n_levels):
            m = pool(m)
            per_level[k].append(m)
    return [torch.stack(lv, 0) for lv in per_level]


def merged_labels(outs: torch.Tensor, n_valid: torch.Tensor) -> torch.Tensor:
    """outs [B,O,HW] -> uint8 labels [B,HW] (evaluator.py:139-145); a video without valid objects gives zeros
    (the reference would raise on the empty max)."""
    B, O, HW = outs.shape
    lab = torch.zeros(B, HW, dtype=torch.uint8)
    for b in range(B):
        n = int(n_valid[b])
        if n == 0:
            continue
        refine_mask = outs[b, :n].view(-1, HW)
        refine_bg = 1 - refine_mask.max(0)[0]
        refine_fbg = torch.cat([refine_bg.view(1, HW), refine_mask], dim=0)
        lab[b] = refine_fbg.max(0)[1].to(torch.uint8)
    return lab


def hard_iou_mean(y_mask: torch.Tensor, pred: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """trainer.py:189-196: mean over the valid templates of the hard IoU between ground truth and prediction."""
    B, O = valid.shape
    n = valid.sum()
    iou = rowwise_binary_iou(y_mask.reshape(B * O, -1), pred.reshape(B * O, -1)).view(B, O) * valid.float()
    return iou.sum() / (n + 1e-6) if n > 0 else torch.zeros_like(iou).sum()


# ---- masker.py:91-173 --------------------------------------------------------------------------------------------
def _expand_boxes(boxes: torch.Tensor, scale: float) -> torch.Tensor:
    w_half = (boxes[:, 2] - boxes[:, 0]) * .5
    h_half = (boxes[:, 3] - boxes[:, 1]) * .5
    x_c = (boxes[:, 2] + boxes[:, 0]) * .5
    y_c = (boxes[:, 3] + boxes[:, 1]) * .5
    w_half = w_half * scale
    h_half = h_half * scale
    out = torch.zeros_like(boxes)
    out[:, 0] = x_c - w_half
    out[:, 2] = x_c + w_half
    out[:, 1] = y_c - h_half
    out[:, 3] = y_c + h_half
    return out


def paste_one(mask: torch.Tensor, box: torch.Tensor, im_h: int, im_w: int, thresh: float = 0.5, padding: int = 1):
    """mask [M,M] soft, box [4] xyxy -> (im_mask [im_h,im_w], tight box int64 [4])  (masker.py:120-173)."""
    M = mask.shape[-1]
    scale = float(M + 2 * padding) / M
    padded = mask.new_zeros((1, 1, M + 2 * padding, M + 2 * padding))
    padded[0, 0, padding:-padding, padding:-padding] = mask
    bx = _expand_boxes(box[None].float(), scale)[0].to(torch.int32)
    x0b, y0b, x1b, y1b = (int(v) for v in bx)
    w = max(x1b - x0b + 1, 1)
    h = max(y1b - y0b + 1, 1)
    res = F.interpolate(padded.float(), size=(h, w), mode="bilinear", align_corners=False)[0, 0]
    im = mask.new_zeros((im_h, im_w), dtype=torch.float32)
    x_0, x_1 = max(x0b, 0), min(x1b + 1, im_w)
    y_0, y_1 = max(y0b, 0), min(y1b + 1, im_h)
    if x_1 > x_0 and y_1 > y_0:          # a box entirely outside the image pastes nothing (the reference would raise)
        im[y_0:y_1, x_0:x_1] = res[(y_0 - y0b):(y_1 - y0b), (x_0 - x0b):(x_1 - x0b)]
    inds = (im > thresh).nonzero()
    if inds.shape[0] < 1:
        tight = torch.tensor([0, 0, im_h, im_w])              # masker.py:160 (rows, cols in that order, as upstream)
    else:
        tight = torch.tensor([inds[:, 1].min().item(), inds[:, 0].min().item(), inds[:, 1].max().item(),
                              inds[:, 0].max().item()])
    return im, tight


def paste_masks(masks: torch.Tensor, boxes: torch.Tensor, im_h: int, im_w: int, thresh: float = 0.5, padding: int = 1):
    """masks [N,1,M,M] (or [N,M,M]), boxes [N,4] -> (pasted [N,im_h,im_w], tight boxes int64 [N,4])
    (Masker.forward_single_image, masker.py:181-206)."""
    masks = masks.reshape(masks.shape[0], masks.shape[-2], masks.shape[-1])
    ims, tights = [], []
    for m, b in zip(masks, boxes):
        im, tb = paste_one(m, b, im_h, im_w, thresh, padding)
        ims.append(im)
        tights.append(tb)
    if not ims:
        return masks.new_zeros((0, im_h, im_w)), torch.zeros(0, 4, dtype=torch.int64)
    return torch.stack(ims, 0), torch.stack(tights, 0)


# ---- boxlist_ops.py:15-29 ----------------------------------------------------------------------------------------
def box_nms(boxes: torch.Tensor, scores: torch.Tensor, thresh: float) -> torch.Tensor:
    """Greedy NMS as in maskrcnn_benchmark's nms (legacy +1 widths); returns kept indices, score-descending.
    Ties in score are broken by the lower index first (stable sort)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.int64)
    order = torch.sort(scores.float(), descending=True, stable=True)[1]
    b = boxes.float()[order]
    area = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    removed = torch.zeros(n, dtype=torch.bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(int(order[i]))
        for j in range(i + 1, n):
            if removed[j]:
                continue
            left, right = torch.max(b[i, 0], b[j, 0]), torch.min(b[i, 2], b[j, 2])
            top, bottom = torch.max(b[i, 1], b[j, 1]), torch.min(b[i, 3], b[j, 3])
            iw = torch.clamp(right - left + 1, min=0.0)
            ih = torch.clamp(bottom - top + 1, min=0.0)
            inter = iw * ih
            if inter / (area[i] + area[j] - inter) > thresh:
                removed[j] = True
    return torch.tensor(keep, dtype=torch.int64)


def filter_results(boxes: torch.Tensor, scores: torch.Tensor, nms_thresh: float = 0.8, max_proposals: int = 0) -> torch.Tensor:
    """boxlist_ops.py:15-29: NMS on the tight boxes, then keep at most max_proposals."""
    keep = box_nms(boxes, scores, nms_thresh)
    if 0 < max_proposals < boxes.shape[0]:
        keep = keep[:max_proposals]
    return keep
