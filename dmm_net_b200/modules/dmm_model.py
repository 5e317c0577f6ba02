"""Batched drop-in for dmm/modules/dmm_model.py (the caller of the matching layer).

Same class name, constructor and method signatures/returns as the reference ``DMM_Model``; what changes is HOW one
frame of a batch of videos is matched: the reference loops over videos in Python with an ``.item()`` sync per video
(dmm_model.py:117) and three 0/1-matrix ``torch.mm`` select/scatter passes (:133-135,:156); here all videos of the
batch go through the kernels in one launch each, valid-template counts stay on the device (``n_tmpl``), the
scatter of the O valid rows into the F=maxseqlen output slots is the apply kernel's ``row_map``, and the per-video
proposal mask tensors are read in place through a device pointer table (no ``torch.stack`` copy).
"""
import torch
import torch.nn as nn

from .. import ops
from ..utils.checker import CHECK4D, CHECKEQ
from .feature_extractor import make_roi_mask_feature_extractor
from .match_model import MatchModel


def _stack_pad(tensors, width, fill=0.0):
    """list of [n_b, ...] -> [B, width, ...] (one copy; skipped when every n_b == width and inputs already form a batch)."""
    if all(t.shape[0] == width for t in tensors):
        return torch.stack(tensors, 0)
    out = tensors[0].new_full((len(tensors), width) + tuple(tensors[0].shape[1:]), fill)
    for b, t in enumerate(tensors):
        out[b, :t.shape[0]] = t
    return out


class DMM_Model(nn.Module):
    r""" container for all DMM modules: match_layer, feature_extractor (reference dmm_model.py:11-20) """

    def __init__(self, cfgs, is_test=0):
        super(DMM_Model, self).__init__()
        self.match_layer = MatchModel(cfgs, is_test)
        self.feature_extractor = make_roi_mask_feature_extractor()
        self.match_algo = cfgs['matching']['algo']

    # ------------------------------------------------------------------------------------------------------
    def fill_template_dict(self, args, proposals, features, y_mask, tplt_valid_batch):
        """First frame: pooled features of the ground-truth boxes become the templates (reference :22-46)."""
        backbone_feature = features['backbone_feature']
        refine_input_feat = features['refine_input_feat']
        boxes_per_image = [len(box) for box in proposals]
        result_alllevel = self.feature_extractor(backbone_feature, proposals)
        feats = result_alllevel.split(boxes_per_image, dim=0)
        tplt_dict = {}
        for b, feat in enumerate(feats):
            tplt_dict[b] = {'feat': [feat], 'refine_input_feat': [tuple([f[b] for f in refine_input_feat])]}
        return tplt_dict

    def prepare_tplt_feature(self, tplt_valid_vid, tplt_dict, bid):
        """Valid template features of one video: rows :O of diag(valid) @ feat (reference :144-158); kept for API
        compatibility -- the batched path does this for all videos at once in ``_gather``."""
        O = int(tplt_valid_vid.sum().item())
        OF_matrix = torch.diag(tplt_valid_vid).float()[:O, :]
        tplt_feat = tplt_dict[bid]['feat']
        assert (type(tplt_feat) == list)
        self.tplt_feat_shape = tplt_feat[0].shape
        return [OF_matrix @ t for t in tplt_feat], OF_matrix

    # ------------------------------------------------------------------------------------------------------
    def _gather(self, proposals, backbone_feature, tplt_dict, tplt_valid_batch, skip=None):
        B = len(proposals)
        boxes_per_image = [len(box) for box in proposals]
        Pmax = max(max(boxes_per_image), 1)
        pooled = self.feature_extractor(backbone_feature, proposals).split(boxes_per_image, dim=0)
        prop_feat = _stack_pad(list(pooled), Pmax)
        # per-video proposal masks stay where they are: the kernels read them through a device pointer table
        prop_mask = ops.RaggedMasks([p.get_field('mask').squeeze(1) for p in proposals])
        prop_score = _stack_pad([p.get_field('objectness') if 'objectness' in p.fields() else p.get_field('scores')
                                 for p in proposals], Pmax)
        dev = prop_mask.device
        H, W = prop_mask.shape_hw[-2:]
        valid = tplt_valid_batch.to(dev).float().view(B, -1)                            # [B,F] 0/1
        Fm = valid.shape[1]
        n_tmpl = valid.sum(1).round().to(torch.int32)                                  # stays on the device
        if skip is not None:
            n_tmpl = torch.where(skip.to(dev), torch.zeros_like(n_tmpl), n_tmpl)
        # reference semantics: rows :O of diag(valid) -> row i is valid[i]*e_i (dmm_model.py:152-154)
        tmpl_feat = self._stacked_templates(tplt_dict, B) * valid[:, None, :, None]                  # [B,T,F,D]
        ar = torch.arange(Fm, device=dev, dtype=torch.int32)[None, :].expand(B, -1)
        row_map = torch.where(valid > 0, ar, torch.full_like(ar, -1)).contiguous()      # scatter fused into K4
        n_prop = torch.tensor(boxes_per_image, dtype=torch.int32, device=dev)
        return prop_feat, prop_mask, prop_score, tmpl_feat, n_prop, n_tmpl, row_map, Fm, (H, W)

    def _match(self, proposals, backbone_feature, mask_last_occurence, tplt_dict, tplt_valid_batch, targets, skip=None):
        B, F, H, W = CHECK4D(mask_last_occurence)
        CHECKEQ(len(proposals), B)
        prop_feat, prop_mask, prop_score, tmpl_feat, n_prop, n_tmpl, row_map, Fm, hw = self._gather(
            proposals, backbone_feature, tplt_dict, tplt_valid_batch, skip)
        CHECKEQ(Fm, F)
        CHECKEQ(tuple(hw), (H, W))
        out = self.match_layer.forward_many(prop_feat, prop_mask, tmpl_feat, mask_last_occurence, prop_score, targets,
                                            n_prop=n_prop, n_tmpl=n_tmpl, row_map=row_map, out_rows=F)
        output_mask = out['full_outmask']                                              # [B,F,H,W], zero rows where invalid
        empty = (n_tmpl == 0).view(B, 1, 1, 1)
        # videos without templates keep their previous masks (reference :66-69,:118-122)
        out_mask_last = torch.where(empty, mask_last_occurence.to(output_mask.dtype), output_mask)
        return output_mask, out_mask_last, out, n_tmpl

    def inference(self, infos, proposals, backbone_feature, mask_last_occurence, tplt_dict, target=None):
        """reference :48-86 -> (output_mask [B,F,H,W], tplt_dict, match_loss [], out_mask_last [B,F,H,W])"""
        extra = infos['extra_frame']
        skip = torch.as_tensor(extra).bool().view(-1) if extra is not None else None
        output_mask, out_mask_last, _, _ = self._match(proposals, backbone_feature, mask_last_occurence, tplt_dict,
                                                       infos['valid'], None if target is None else torch.stack(list(target), 0)
                                                       if not torch.is_tensor(target) else target, skip)
        return output_mask, tplt_dict, [], out_mask_last

    def forward(self, args, proposals, backbone_feature, mask_last_occurence, tplt_dict, tplt_valid_batch, targets):
        """reference :88-142 -> (output_mask [B,F,H,W], tplt_dict, match_loss list (one scalar per video), out_mask_last)"""
        assert (targets is not None)
        output_mask, out_mask_last, out, n_tmpl = self._match(proposals, backbone_feature, mask_last_occurence, tplt_dict,
                                                              tplt_valid_batch, targets)
        loss = out['cost_loss'] * (n_tmpl > 0).float()       # videos without templates contribute prop_feat.sum()*0
        match_loss = list(loss.unbind(0))
        return output_mask, tplt_dict, match_loss, out_mask_last

    def _stacked_templates(self, tplt_dict, B):
        """[B,T,F,D] stack of the per-video template features; the dict does not change between the frames of a clip, so
        the stack is cached on the identity of its tensors instead of being rebuilt (B small copies) every frame."""
        T = len(tplt_dict[0]['feat'])
        if torch.is_grad_enabled() and any(tplt_dict[b]['feat'][t].requires_grad for b in range(B) for t in range(T)):
            # training: every frame gets its own stack node in the autograd graph, like the reference's per-frame rebuild
            return torch.stack([torch.stack([tplt_dict[b]['feat'][t] for t in range(T)], 0) for b in range(B)], 0)
        # identity AND content version of every tensor: an in-place update (copy_, mul_, a momentum update) keeps id()
        key = tuple((id(x), x.data_ptr(), x._version) for x in (tplt_dict[b]['feat'][t] for b in range(B) for t in range(T)))
        cached = getattr(self, '_tmpl_cache', None)
        if cached is None or cached[0] != key:
            stacked = torch.stack([torch.stack([tplt_dict[b]['feat'][t] for t in range(T)], 0) for b in range(B)], 0)
            cached = (key, stacked, [tplt_dict[b]['feat'] for b in range(B)])      # keeps the tensors (and their ids) alive
            self._tmpl_cache = cached
        return cached[1]

    # ------------------------------------------------------------------------------------------------------
    def inference_lazy(self, infos, detections, backbone_feature, mask_last_occurence, tplt_dict, nms_thresh=0.8,
                       max_proposals=50, mask_threshold=0.5, padding=1):
        """``inference`` fed with the mask head's raw outputs instead of pasted masks (SURVEY.md 8f-2 + 8f-4 together).

        The reference pastes every detection into a full-resolution soft mask (masker.py), runs NMS on the tight boxes
        (boxlist_ops.py), then matches; per frame that writes P*HW*4 bytes, reads them for the IoU and reads the selected
        ones again for ``torch.mm``.  The IoU only needs one bit per pixel and the apply only the <= O selected
        detections, so here the soft masks are never materialised:
          K8 (bits + tight boxes only) -> K9 NMS -> K5 features of the kept boxes -> K2 -> K1 on packed rows -> K3 ->
          K10 pastes the selected detections, scaled, straight into the output rows.
        No host synchronisation anywhere (the keep lists stay on the device as an index table).  Outputs are
        bit-identical to ``Masker`` -> ``filter_results`` -> ``inference`` (tests/test_gpu_container.py).

        detections: list (one per video) of BoxList-likes with ``bbox`` [n,4] (detector boxes), field ``mask`` [n,1,M,M]
        (mask-head probabilities, NOT pasted) and ``scores`` / ``objectness`` [n].
        Returns (output_mask [B,F,H,W], tplt_dict, [], out_mask_last, keep) -- ``keep`` = (index table [B,P], n_prop [B])."""
        extra = infos.get('extra_frame')
        skip = torch.as_tensor(extra).bool().view(-1) if extra is not None else None
        with torch.no_grad():
            output_mask, out_mask_last, _, keep = self._lazy(detections, backbone_feature, mask_last_occurence, tplt_dict,
                                                            infos['valid'], None, skip, nms_thresh, max_proposals, mask_threshold,
                                                            padding)
        return output_mask, tplt_dict, [], out_mask_last, keep

    def forward_lazy(self, args, detections, backbone_feature, mask_last_occurence, tplt_dict, tplt_valid_batch, targets,
                     nms_thresh=0.8, max_proposals=50, mask_threshold=0.5, padding=1):
        """``forward`` (training: match loss + gradients) through the lazy pipeline: the P pasted proposal masks of
        masker.py:181-206 are never materialised in training either.  Gradients flow through K10's backward into the
        assignment, through the solver and the cosine into the pooled features and the backbone maps; the detections'
        mask-head outputs are constants (offline proposals in the reference).
        Returns (output_mask [B,F,H,W], tplt_dict, match_loss list, out_mask_last, keep)."""
        assert (targets is not None)
        output_mask, out_mask_last, cost_loss, keep = self._lazy(detections, backbone_feature, mask_last_occurence, tplt_dict,
                                                                 tplt_valid_batch, targets, None, nms_thresh, max_proposals,
                                                                 mask_threshold, padding)
        return output_mask, tplt_dict, list(cost_loss.unbind(0)), out_mask_last, keep

    def _lazy(self, detections, backbone_feature, mask_last_occurence, tplt_dict, tplt_valid_batch, targets, skip, nms_thresh,
              max_proposals, mask_threshold, padding):
        B, F, H, W = CHECK4D(mask_last_occurence)
        CHECKEQ(len(detections), B)
        dev = mask_last_occurence.device
        counts = [len(d) for d in detections]
        n_max = max(max(counts), 1)
        m_all = torch.cat([d.get_field('mask').reshape(len(d), d.get_field('mask').shape[-2], d.get_field('mask').shape[-1])
                           for d in detections], 0)
        b_all = torch.cat([d.bbox for d in detections], 0).to(dev)
        with torch.no_grad():
            pasted = ops.paste_masks(m_all, b_all, H, W, mask_threshold, padding, want_pasted=False, want_bits=True)   # K8
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        sc_all = torch.cat([(d.get_field('objectness') if 'objectness' in d.fields() else d.get_field('scores')).float()
                            for d in detections], 0)
        if all(c == n_max for c in counts):                      # equal counts: the flat lists already are the [B, n] tables
            tight = pasted['tight'].float().view(B, n_max, 4)
            score = sc_all.view(B, n_max)
        else:                                                    # ragged: ONE scatter through a host-built index (no sync)
            flat = torch.tensor([b * n_max + j for b, c in enumerate(counts) for j in range(c)], dtype=torch.long).to(dev)
            tight = torch.zeros(B * n_max, 4, device=dev).index_copy_(0, flat, pasted['tight'].float()).view(B, n_max, 4)
            score = torch.zeros(B * n_max, device=dev).index_copy_(0, flat, sc_all).view(B, n_max)
        cnt = torch.tensor(counts, dtype=torch.int32, device=dev)
        keep, n_keep = ops.box_nms(tight, score.detach(), nms_thresh, max_proposals, cnt)                           # K9
        P = n_max if max_proposals <= 0 else min(max_proposals, n_max)
        keep = keep[:, :P]
        n_prop = n_keep.clamp(max=P)
        kept = keep >= 0
        local = keep.clamp(min=0)
        src_index = torch.where(kept, local + torch.tensor(offs[:-1], device=dev)[:, None], torch.full_like(local, -1)).to(torch.int32)
        gidx = src_index.clamp(min=0).long()
        rois = torch.cat([torch.arange(B, device=dev, dtype=torch.float32)[:, None, None].expand(B, P, 1),
                          pasted['tight'][gidx].float()], 2).view(B * P, 5)
        prop_feat = self.feature_extractor.pool_rois(backbone_feature, rois).view(B, P, -1)                          # K5
        prop_bits = pasted['bits'][gidx]                                                                             # [B,P,words]
        prop_score = torch.gather(score, 1, local) * kept
        valid = tplt_valid_batch.to(dev).float().view(B, -1)
        CHECKEQ(valid.shape[1], F)
        n_tmpl = valid.sum(1).round().to(torch.int32)
        if skip is not None:
            n_tmpl = torch.where(skip.to(dev), torch.zeros_like(n_tmpl), n_tmpl)
        tmpl_feat = self._stacked_templates(tplt_dict, B) * valid[:, None, :, None]
        ar = torch.arange(F, device=dev, dtype=torch.int32)[None, :].expand(B, -1)
        row_map = torch.where(valid > 0, ar, torch.full_like(ar, -1)).contiguous()
        layer = self.match_layer
        w = float(layer.cfgs['score_weight'])
        cos = ops.cosine_pairwise(tmpl_feat, prop_feat, n_prop, n_tmpl)                                              # K2
        with torch.no_grad():
            tmpl_bits = ops.pack_masks(mask_last_occurence.float(), mask_dims=2)                                     # [B,F,words]
            tgt_bits = None if targets is None else ops.pack_masks(targets.float(), mask_dims=2)
        cost_loss = None
        if targets is None:
            r = ops.mask_iou_pairwise_packed(prop_bits.contiguous(), tmpl_bits, None, n_prop, n_tmpl, cos=cos.detach(),
                                             w_cos=1 - w, w_iou=w)                                                   # K1 packed
            sim = r['sim']
        else:
            r = ops.mask_iou_pairwise_packed(prop_bits.contiguous(), tmpl_bits, tgt_bits, n_prop, n_tmpl)            # both sets, one pass
            sim = cos * (1 - w) + r['iou'] * w                                                                       # autograd sees cos
            gt = ops.relax_solve(r['iou2'], None, n_prop, n_tmpl, 0, 0, 0.0, True, False, True)[0]                   # greedy one-hot
            ok = (torch.arange(P, device=dev)[None, None, :] < n_prop[:, None, None]) & \
                 (torch.arange(F, device=dev)[None, :, None] < n_tmpl[:, None, None])
            cost_loss = (((cos - gt) ** 2) * ok).sum(dim=(1, 2)) / (n_prop * n_tmpl).clamp(min=1).float()
            cost_loss = cost_loss * (n_tmpl > 0).float()
        _, Bm, _, _, _, logic, _ = ops.relax_solve(sim, prop_score, n_prop, n_tmpl, layer.max_iter, layer.proj_iter,
                                                   layer.relax_lr, True, True, bool(layer.is_test))                  # K3
        output_mask = ops.paste_apply(Bm, m_all, b_all, src_index, H, W, n_prop, n_tmpl, row_map, F, True, padding,
                                      logic=logic)                                                                   # K10
        empty = (n_tmpl == 0).view(B, 1, 1, 1)
        out_mask_last = torch.where(empty, mask_last_occurence.to(output_mask.dtype), output_mask)
        return output_mask, out_mask_last, cost_loss, (src_index, n_prop)
