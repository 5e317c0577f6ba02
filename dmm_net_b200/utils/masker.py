"""Drop-in for the paste part of dmm/utils/masker.py: ``Masker`` projects the fixed-size proposal masks into the image
at their boxes.  The reference loops over proposals in Python (zero-pad, int box, F.interpolate, slice-assign, nonzero +
four .item() syncs each, masker.py:120-206); here all proposals of all images go through ONE K8 launch that writes the
pasted soft masks, and on request their thresholded bit rows for the packed K1 entry."""
import torch

from .. import ops


def paste_mask_in_image(mask, box, im_h, im_w, thresh=0.5, padding=1):
    """mask [M,M], box [4] -> (im_mask [im_h,im_w], tight box int64 [4])   (masker.py:120-150)."""
    r = ops.paste_masks(mask[None], box[None], im_h, im_w, thresh, padding)
    return r["pasted"][0], r["tight"][0]


class Masker(object):
    """Projects a set of masks in an image on the locations specified by the bounding boxes (masker.py:169-230)."""

    def __init__(self, threshold, padding=1):
        self.threshold = threshold
        self.padding = padding

    def forward_single_image(self, masks, boxes):
        """masks [Nbox,1,M,M]; boxes: BoxList-like with ``bbox`` [Nbox,4] xyxy and ``size`` = (im_w, im_h).
        Returns (pasted [Nbox,1,im_h,im_w], tight boxes [Nbox,4] int64)."""
        res, resb = self([masks], [boxes])
        return res[0], resb[0]

    def __call__(self, masks, boxes, want_bits=False):
        if not isinstance(boxes, (list, tuple)):
            boxes = [boxes]
        if torch.is_tensor(masks):
            masks = [masks]
        assert len(boxes) == len(masks), "Masks and boxes should have the same length."
        for mask, box in zip(masks, boxes):
            assert mask.shape[0] == len(box), "Number of objects should be the same."
        # reference masker.py:184 converts every BoxList to xyxy before pasting
        boxes = [b.convert("xyxy") if hasattr(b, "convert") else b for b in boxes]
        for b in boxes:
            assert getattr(b, "mode", "xyxy") == "xyxy", "Masker pastes xyxy boxes (BoxList.convert('xyxy') is missing)"
            assert b.size is not None and len(tuple(b.size)) == 2, \
                "Masker needs the image size: BoxList(bbox, image_size=(im_w, im_h))"
        sizes = {tuple(b.size) for b in boxes}
        results, results_box, results_bits = [None] * len(boxes), [None] * len(boxes), [None] * len(boxes)
        for size in sizes:                                     # images of one batch share a size: one launch
            im_w, im_h = size
            ids = [i for i, b in enumerate(boxes) if tuple(b.size) == size]
            m = torch.cat([masks[i].reshape(masks[i].shape[0], masks[i].shape[-2], masks[i].shape[-1]) for i in ids], 0)
            bb = torch.cat([boxes[i].bbox for i in ids], 0)
            if m.shape[0] == 0:
                for i in ids:
                    results[i] = masks[i].new_empty((0, 1, masks[i].shape[-2], masks[i].shape[-1]))   # masker.py:201
                    results_box[i] = boxes[i].bbox
                continue
            r = ops.paste_masks(m, bb.to(m.device), im_h, im_w, self.threshold, self.padding, want_bits=want_bits)
            off = 0
            for i in ids:
                n = len(boxes[i])
                if n == 0:
                    results[i] = masks[i].new_empty((0, 1, masks[i].shape[-2], masks[i].shape[-1]))
                    results_box[i] = boxes[i].bbox
                else:
                    results[i] = r["pasted"][off:off + n][:, None]
                    results_box[i] = r["tight"][off:off + n]
                    if want_bits:
                        results_bits[i] = r["bits"][off:off + n]
                off += n
        if want_bits:
            return results, results_box, results_bits
        return results, results_box
