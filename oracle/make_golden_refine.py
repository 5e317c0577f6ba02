#!/usr/bin/env python
"""Generate tests/golden/{refine,paste}_*.npz from the reference's OWN code (build container only).

    python oracle/make_golden_refine.py

* ``refine_pyramid_*``: the reference's decoder-input lines (dmm/modules/trainer.py:256-263) are READ from
  /root/reference at run time, dedented and exec'ed per object with seeded inputs (they sit inside ``Trainer.refine``,
  which cannot be imported: it pulls in the un-vendored maskrcnn_benchmark).  Also stores autograd gradients through
  those lines.  Nothing of the reference is copied into this repository.
* ``refine_labels_*``: same for the merged-label lines dmm/modules/evaluator.py:139-143.
* ``paste_*``: ``dmm.utils.masker.paste_mask_in_image`` imported from the reference with the two maskrcnn_benchmark
  names it needs stubbed (``interpolate`` -> ``torch.nn.functional.interpolate``, which is what that wrapper calls for
  non-empty inputs; ``BoxList`` -> placeholder class, unused by the function).
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch
from torch import nn  # noqa: F401  (name used by the exec'ed reference lines)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)


def ref_lines(rel, first, last):
    with open(os.path.join(REF, rel)) as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def npy(t):
    return t.detach().cpu().numpy()


def soft_masks(g, B, O, H, W, fill=0.35):
    """soft masks with exact zeros outside a random box (what paste_mask_in_image / sigmoid heads deliver) and ties"""
    m = torch.zeros(B, O, H, W)
    for b in range(B):
        for o in range(O):
            if torch.rand((), generator=g) < 0.15:
                continue                                               # an all-zero plane
            y0 = int(torch.randint(0, max(H // 2, 1), (), generator=g)); x0 = int(torch.randint(0, max(W // 2, 1), (), generator=g))
            y1 = min(H, y0 + 1 + int(torch.randint(0, H, (), generator=g))); x1 = min(W, x0 + 1 + int(torch.randint(0, W, (), generator=g)))
            blob = torch.rand(y1 - y0, x1 - x0, generator=g)
            blob = torch.where(blob < fill, torch.zeros(()), (blob * 8).round() / 8)   # quantised: many exact ties
            m[b, o, y0:y1, x0:x1] = blob
    return m


def pyramid_case(name, B, O, H, W, L, seed):
    g = torch.Generator().manual_seed(seed)
    prev = soft_masks(g, B, O, H, W).view(B, O, H * W).requires_grad_(True)
    ref = soft_masks(g, B, O, H, W).view(B, O, H * W).requires_grad_(True)
    init = soft_masks(g, B, O, H, W).requires_grad_(True)
    src = ref_lines("dmm/modules/trainer.py", 256, 263)
    levels = [[] for _ in range(L)]
    for t in range(O):
        ns = {"nn": nn, "torch": torch, "prev_mask": prev, "ref_mask": ref, "init_pred_inst": init, "obj_index": t,
              "B": B, "H": H, "W": W, "feats": [None] * L}
        exec(src, ns)
        for k in range(L):
            levels[k].append(ns["mask_lstm"][k])
    levels = [torch.stack(lv, 0) for lv in levels]                      # [O,B,3,hk,wk]
    ws = [torch.rand(lv.shape, generator=g) for lv in levels]
    sum((lv * w).sum() for lv, w in zip(levels, ws)).backward()
    out = {"prev": npy(prev), "ref": npy(ref), "init": npy(init), "meta": np.array([B, O, H, W, L], np.int64),
           "g_prev": npy(prev.grad), "g_ref": npy(ref.grad), "g_init": npy(init.grad)}
    for k in range(L):
        out[f"level{k}"] = npy(levels[k])
        out[f"w{k}"] = npy(ws[k])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, [tuple(lv.shape) for lv in levels])


def labels_case(name, B, O, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    outs = torch.sigmoid(4 * soft_masks(g, B, O, H, W, fill=0.5) - 2).view(B, O, H * W)
    outs[:, :, : H * W // 7] = 0.5                                      # exact ties between background and objects
    n_valid = torch.randint(1, O + 1, (B,), generator=g)
    tplt_valid_batch = (torch.arange(O)[None] < n_valid[:, None]).long()
    src = ref_lines("dmm/modules/evaluator.py", 139, 143)
    labs = []
    for b in range(B):
        ns = {"torch": torch, "outs": outs, "tplt_valid_batch": tplt_valid_batch, "b": b, "H": H, "W": W}
        exec(src, ns)
        labs.append(ns["max_i"].view(-1))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), outs=npy(outs), n_valid=npy(n_valid),
                        label=npy(torch.stack(labs, 0)).astype(np.uint8), meta=np.array([B, O, H, W], np.int64))
    print(name, tuple(outs.shape))


def import_ref_masker():
    import torch.nn.functional as F
    mb = types.ModuleType("maskrcnn_benchmark")
    layers = types.ModuleType("maskrcnn_benchmark.layers")
    misc = types.ModuleType("maskrcnn_benchmark.layers.misc")
    misc.interpolate = F.interpolate
    structures = types.ModuleType("maskrcnn_benchmark.structures")
    bb = types.ModuleType("maskrcnn_benchmark.structures.bounding_box")
    bb.BoxList = type("BoxList", (), {})
    for n, m in (("maskrcnn_benchmark", mb), ("maskrcnn_benchmark.layers", layers), ("maskrcnn_benchmark.layers.misc", misc),
                 ("maskrcnn_benchmark.structures", structures), ("maskrcnn_benchmark.structures.bounding_box", bb)):
        sys.modules.setdefault(n, m)
    sys.path.insert(0, REF)
    from dmm.utils import masker
    return masker


def paste_case(name, masker, N, M, im_h, im_w, seed, thresh=0.5):
    g = torch.Generator().manual_seed(seed)
    masks = torch.sigmoid(3 * torch.randn(N, 1, M, M, generator=g))
    cx = torch.rand(N, generator=g) * im_w
    cy = torch.rand(N, generator=g) * im_h
    bw = torch.rand(N, generator=g) * im_w * 0.6 + 1
    bh = torch.rand(N, generator=g) * im_h * 0.6 + 1
    boxes = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    boxes[:, 0::2].clamp_(0, im_w - 1)                                  # BoxList.clip_to_image
    boxes[:, 1::2].clamp_(0, im_h - 1)
    if N > 3:
        boxes[1] = torch.tensor([3.2, 4.7, 3.9, 4.9])                   # sub-pixel box -> 1-2 px paste
        boxes[2] = torch.tensor([0.0, 0.0, im_w - 1.0, im_h - 1.0])     # whole image: expanded box sticks out on all sides
        masks[3] = 0.1                                                  # nothing above the threshold -> fallback tight box
    ims, tights = [], []
    for m, b in zip(masks, boxes):
        im, tb = masker.paste_mask_in_image(m[0], b, im_h, im_w, thresh, 1)
        ims.append(im)
        tights.append(tb.long())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), masks=npy(masks), boxes=npy(boxes), pasted=npy(torch.stack(ims)),
                        tight=npy(torch.stack(tights)), meta=np.array([N, M, im_h, im_w], np.int64), thresh=np.float64(thresh))
    print(name, N, (im_h, im_w))


def main():
    os.makedirs(OUT, exist_ok=True)
    pyramid_case("refine_pyramid_a", 2, 3, 64, 112, 4, 11)              # multiples of the 64-px tile and not
    pyramid_case("refine_pyramid_odd", 1, 2, 37, 53, 4, 12)             # odd sizes: clipped windows on both edges
    pyramid_case("refine_pyramid_tall", 2, 1, 130, 66, 5, 13)           # 5 levels, just past a tile boundary
    labels_case("refine_labels_a", 3, 5, 32, 56, 21)
    labels_case("refine_labels_odd", 2, 3, 17, 23, 22)
    masker = import_ref_masker()
    paste_case("paste_a", masker, 12, 28, 64, 112, 31)
    paste_case("paste_odd", masker, 7, 28, 45, 83, 32)
    print("written to", OUT)


if __name__ == "__main__":
    main()
