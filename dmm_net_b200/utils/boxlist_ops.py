"""Drop-in for dmm/utils/boxlist_ops.py:15-29 (``filter_results``): NMS on the tight boxes of every frame in one K9 launch
instead of one maskrcnn_benchmark ``nms`` call (and one host sync) per frame."""
import torch

from .. import ops


def filter_results(boxlists, nms_thresh=0.8, max_proposals=0, score_field="scores"):
    """list of BoxList-likes (``bbox`` [n,4], field ``score_field``; must support ``boxlist[keep]``) -> the same list,
    each entry indexed by its kept proposals in score order."""
    F_ = len(boxlists)
    if F_ == 0:
        return boxlists
    n_max = max(len(b) for b in boxlists)
    if n_max == 0:
        return boxlists
    dev = boxlists[0].bbox.device
    boxes = torch.zeros(F_, n_max, 4, device=dev)
    scores = torch.zeros(F_, n_max, device=dev)
    for f, b in enumerate(boxlists):
        boxes[f, :len(b)] = b.bbox.float()
        scores[f, :len(b)] = b.get_field(score_field).float()
    counts = torch.tensor([len(b) for b in boxlists], dtype=torch.int32, device=dev)
    keep, n_keep = ops.box_nms(boxes, scores, nms_thresh, max_proposals, counts)
    n_keep = n_keep.tolist()                                   # one sync for the whole batch
    for f in range(F_):
        boxlists[f] = boxlists[f][keep[f, :n_keep[f]]]
    return boxlists
