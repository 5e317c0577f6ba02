"""Minimal stand-in for maskrcnn_benchmark.structures.bounding_box.BoxList (un-vendored dependency of the
reference): just what the matching path touches -- ``bbox`` [n,4] xyxy, ``get_field/add_field/fields`` and ``len``
(dmm/modules/dmm_model.py:59-71,106-123; dmm/modules/feature_extractor.py:32-37) plus ``boxlist[keep]`` for
``filter_results`` (dmm/utils/boxlist_ops.py:28)."""
import torch


class BoxList(object):
    def __init__(self, bbox, image_size=None, mode="xyxy"):
        bbox = torch.as_tensor(bbox, dtype=torch.float32)
        assert bbox.dim() == 2 and bbox.shape[-1] == 4, bbox.shape
        assert mode in ("xyxy", "xywh"), mode
        self.bbox, self.size, self.mode = bbox, image_size, mode
        self.extra_fields = {}

    def convert(self, mode):
        """xyxy <-> xywh with the legacy +1 pixel convention of maskrcnn_benchmark's BoxList.convert (TO_REMOVE = 1)."""
        assert mode in ("xyxy", "xywh"), mode
        if mode == self.mode:
            return self
        x0, y0, a, b = self.bbox.unbind(-1)
        if mode == "xyxy":                                       # from xywh
            bbox = torch.stack([x0, y0, x0 + (a - 1).clamp(min=0), y0 + (b - 1).clamp(min=0)], -1)
        else:                                                    # xyxy -> xywh
            bbox = torch.stack([x0, y0, a - x0 + 1, b - y0 + 1], -1)
        out = BoxList(bbox, self.size, mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v)
        return out

    def add_field(self, name, data):
        self.extra_fields[name] = data

    def get_field(self, name):
        return self.extra_fields[name]

    def has_field(self, name):
        return name in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]
