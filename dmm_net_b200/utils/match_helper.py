"""Drop-in for dmm/utils/match_helper.py: same three functions, same signatures, CUDA kernels underneath."""
import torch
import torch.nn.functional as F

from .. import ops
from .checker import CHECK2D, CHECK3D, CHECKEQ


def compute_iou_binary_mask_2D(annotation, segmentation):
    """[N,M], [N,M] -> [N] hard IoU at threshold 0.5, no gradient (reference match_helper.py:9-28).  K1 row-paired."""
    CHECK2D(annotation)
    CHECK2D(segmentation)
    CHECKEQ(annotation.shape, segmentation.shape)
    with torch.no_grad():
        iou = ops.mask_iou_rowwise(annotation.float(), segmentation.float())
    return iou.detach()


def get_cosine_score(query_feature, key_feature, cfgs=None):
    """[O,D] x [P,D] -> [O,P] cosine similarity, differentiable (reference match_helper.py:51-64).  K2."""
    CHECK2D(query_feature)
    CHECK2D(key_feature)
    CHECKEQ(query_feature.shape[1], key_feature.shape[1])
    return ops.cosine_pairwise(query_feature[None, None], key_feature[None])[0]


def compute_matching_loss(proposed_mask, targets, similarity_matrix, cfgs=None):
    """MSE between the feature similarity and the greedy one-hot matching of IoU(proposals, targets)
    (reference match_helper.py:30-49).  K1 + the solver's greedy prologue."""
    CHECK3D(proposed_mask)
    CHECK3D(targets)
    CHECKEQ(proposed_mask.shape[-1], targets.shape[-1])
    with torch.no_grad():
        iou = ops.mask_iou_pairwise(proposed_mask[None].float(), targets[None].float())["iou"]
        gt = ops.relax_solve(iou, None, max_iter=0, proj_iter=0, lr=0.0, negate=True, pad_rule=False)[0][0]
    return F.mse_loss(similarity_matrix, gt)
