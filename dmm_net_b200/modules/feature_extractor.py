"""Drop-in for dmm/modules/feature_extractor.py: per-proposal features = ROIAlign(14x14, sampling 2) on the 4
feature levels followed by the spatial mean, concatenated -> [R, 4*C].  One fused separable-weights kernel (K5)
instead of 4 ROIAlign launches + a [R,4,C,14,14] intermediate + two means."""
import torch
import torch.nn as nn

from .. import ops


class FeatureExtractor(nn.Module):
    """Heads for FPN-style pooled proposal features (reference feature_extractor.py:5-30)."""

    scales = (0.25, 0.125, 0.0625, 0.03125)   # feature_extractor.py:13
    sampling_ratio = 2                          # :14
    resolution = 14                             # :15

    def __init__(self):
        super(FeatureExtractor, self).__init__()
        self.num_levels = len(self.scales)
        self.output_size = (self.resolution, self.resolution)

    def convert_to_roi_format(self, boxes):
        """list of BoxList-likes (``.bbox`` [n,4] xyxy) -> [sum n, 5] rows (image index, x1, y1, x2, y2) (:32-37)."""
        rows = []
        for i, b in enumerate(boxes):
            bb = b.bbox
            ids = torch.full((bb.shape[0], 1), float(i), dtype=bb.dtype, device=bb.device)
            rows.append(torch.cat([ids, bb], dim=1))
        return rows[0] if len(rows) == 1 else torch.cat(rows, dim=0)

    def pool_rois(self, backbone_feature, rois):
        """rois [R,5] = (image index, x1, y1, x2, y2) already in ROI format -> [R, 4*C] (what ``forward`` computes after
        ``convert_to_roi_format``); used by the lazy pipeline, whose kept boxes live in a fixed-size device table."""
        assert len(backbone_feature) == self.num_levels, len(backbone_feature)
        return ops.roi_mean_pool(list(backbone_feature), rois)

    def forward(self, backbone_feature, proposals):
        """backbone_feature: 4 tensors [B,C,H/s,W/s] (s = 4, 8, 16, 32); proposals: list of B BoxList-likes.
        Returns [sum_b len(proposals[b]), 4*C], level-major per ROI like ``result.mean(4).mean(3).view(R,-1)`` (:20-30)."""
        assert len(backbone_feature) == self.num_levels, len(backbone_feature)
        rois = self.convert_to_roi_format(proposals)
        return ops.roi_mean_pool(list(backbone_feature), rois)


def make_roi_mask_feature_extractor():
    return FeatureExtractor()
