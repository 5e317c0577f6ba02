"""Drop-in for dmm/modules/match_model.py: the differentiable mask-matching layer on sm_100a kernels.

Same constructor, same ``forward`` signature and return tuple as the reference ``MatchModel``
(match_model.py:13-47); ``compute_cost_matrix`` / ``match_with_first_frame`` / ``compute_feature_score`` keep
their names and meaning.  ``forward_many`` is the batched entry the B200 design wants: B problems per launch.
"""
import torch
import torch.nn as nn

from .. import ops
from ..utils.checker import CHECK3D, CHECKEQ
from ..utils import match_helper
from ..utils.match_helper import compute_iou_binary_mask_2D  # re-exported: trainer.py:19 imports it from here  # noqa: F401
from .submodules.relax_match import hungarian_matching, relax_matching  # noqa: F401


class MatchModel(nn.Module):
    def __init__(self, cfgs={}, is_test=0):
        super(MatchModel, self).__init__()
        self.cfgs = cfgs
        self.is_test = is_test
        self.match_algo = cfgs['matching']['algo']
        self.max_iter = self.cfgs['relax_max_iter']
        self.proj_iter = self.cfgs['relax_proj_iter']
        self.relax_lr = self.cfgs['relax_learning_rate']
        assert (self.match_algo == 'relax' or self.match_algo == 'hun')

    # ------------------------------------------------------------------------------------------------------
    def forward(self, proposed_feature, proposed_mask, template_feature, mask_last_occurence, proposal_score, targets=None):
        """proposed_feature [P,D], proposed_mask [P,H,W], template_feature list of [O,D], mask_last_occurence [O,H,W],
        proposal_score [P], targets [O,H,W] or None ->
        (full_outmask [O,H,W], match_score [O], det_score [O], full_outmask (same object), {'cost_loss': scalar} or {})."""
        CHECK3D(proposed_mask)
        CHECK3D(mask_last_occurence)
        CHECKEQ(proposal_score.shape[0], proposed_mask.shape[0])
        if self.match_algo == 'hun':
            return self._forward_hungarian(proposed_feature, proposed_mask, template_feature, mask_last_occurence,
                                           proposal_score, targets)
        out = ops.match_batch(proposed_feature[None], proposed_mask[None], torch.stack(list(template_feature), 0)[None],
                              mask_last_occurence[None], proposal_score[None],
                              None if targets is None else targets[None],
                              max_iter=self.max_iter, proj_iter=self.proj_iter, lr=self.relax_lr,
                              score_weight=self.cfgs['score_weight'], is_test=bool(self.is_test))
        match_loss = {}
        if targets is not None:
            match_loss['cost_loss'] = out['cost_loss'][0]
        full_outmask = out['full_outmask'][0]
        return full_outmask, out['match_score'][0], out['det_score'][0], full_outmask, match_loss

    def forward_many(self, proposed_feature, proposed_mask, template_feature, mask_last_occurence, proposal_score,
                     targets=None, n_prop=None, n_tmpl=None, row_map=None, out_rows=None):
        """Batched layer: [B,P,D], [B,P,H,W], [B,O,D] or [B,T,O,D], [B,O,H,W], [B,P], targets [B,O,H,W] or None.
        Returns the dict of ``ops.match_batch`` (full_outmask [B,O,H,W], match_score/det_score [B,O], cost_loss [B]...)."""
        assert self.match_algo == 'relax', "the batched path implements the relaxed solver"
        return ops.match_batch(proposed_feature, proposed_mask, template_feature, mask_last_occurence, proposal_score,
                               targets, max_iter=self.max_iter, proj_iter=self.proj_iter, lr=self.relax_lr,
                               score_weight=self.cfgs['score_weight'], is_test=bool(self.is_test), n_prop=n_prop,
                               n_tmpl=n_tmpl, row_map=row_map, O_out=out_rows)

    def forward_many_host(self, proposed_feature, proposed_mask, template_feature, mask_last_occurence, proposal_score,
                          device="cuda", threads=None, n_prop=None, n_tmpl=None, raw_fraction=None):
        """cost-build + solve for inputs held in HOST memory (CPU fp32 tensors): the host cores bit-pack most masks (only
        bits cross PCIe) while the copy engine DMAs the rest as fp32 (``ops.match_batch_host``).  Returns device tensors
        (sim, R, Bmat, scores...)."""
        assert self.match_algo == 'relax'
        return ops.match_batch_host(proposed_feature, proposed_mask, template_feature, mask_last_occurence, proposal_score,
                                    max_iter=self.max_iter, proj_iter=self.proj_iter, lr=self.relax_lr,
                                    score_weight=self.cfgs['score_weight'], is_test=bool(self.is_test), device=device,
                                    threads=threads, n_prop=n_prop, n_tmpl=n_tmpl, raw_fraction=raw_fraction)

    # ------------------------------------------------------------------------------------------------------
    def compute_cost_matrix(self, features, mask, scores, targets=None):
        """sim = (1-w) * mean_t cos(template_t, proposals) + w * IoU(template masks, proposal masks), plus the optional
        matching loss on the cosine part.  Returns (sim [O,P], n_prop, n_tplt, match_loss)  (reference :49-91)."""
        proposed_feature, template_feature = features['proposed'], features['template']
        proposed_mask, mask_last_occurence = mask['proposed'], mask['template']
        CHECK3D(proposed_mask)
        n_prop = proposed_mask.shape[0]
        n_tplt = template_feature[0].shape[0]
        feature_sim = ops.cosine_pairwise(torch.stack(list(template_feature), 0)[None], proposed_feature[None])[0]
        match_loss = {}
        if targets is not None:
            match_loss['cost_loss'] = match_helper.compute_matching_loss(proposed_mask, targets, feature_sim, self.cfgs)
        CHECKEQ(proposed_mask.shape[-2:], mask_last_occurence.shape[-2:])
        with torch.no_grad():
            iou = ops.mask_iou_pairwise(proposed_mask[None].float(), mask_last_occurence[None].float())['iou'][0]
        w = self.cfgs['score_weight']
        sim_matrix = feature_sim * (1 - w) + iou * w
        return sim_matrix, n_prop, n_tplt, match_loss

    def match_with_first_frame(self, sim_matrix, n_prop, n_tplt, proposed_mask, proposal_score, mask_last_occurence):
        """sim [O,P] -> (full_outmask [O,H,W], match_score [O], det_score [O], logic_mask [O,m], binary_Ridx_matched [O,m])
        with m = P, or O+1 when P <= O (reference :93-148)."""
        H, W = proposed_mask.shape[-2], proposed_mask.shape[-1]
        O, P = sim_matrix.shape
        m = P if P > O else O + 1
        if self.match_algo == 'relax':
            R, Bm, ms, ds, _, logic, _ = ops.relax_solve(sim_matrix[None], proposal_score[None], None, None, self.max_iter,
                                                         self.proj_iter, self.relax_lr, True, True, bool(self.is_test))
        else:
            R, Bm, ms, ds, logic = self._hungarian_head(sim_matrix, proposal_score)
        full_outmask = ops.assign_apply(Bm, proposed_mask.float()[None], logic)[0].view(-1, H, W)
        return full_outmask, ms[0], ds[0], logic[0, :, :m], Bm[0, :, :m]

    def compute_feature_score(self, key_feature, query_feature):
        return match_helper.get_cosine_score(query_feature, key_feature, self.cfgs)

    # ------------------------------------------------------------------------------------------------------
    def _hungarian_head(self, sim_matrix, proposal_score):
        """algo == 'hun': SciPy assignment on the host (as in the reference), head arithmetic on the tiny [O,m] matrix."""
        O, P = sim_matrix.shape
        m = P if P > O else O + 1
        sim_pad = sim_matrix.new_zeros((O, m))
        sim_pad[:, :P] = sim_matrix
        R, _, _, _ = hungarian_matching(-sim_pad)
        maxv, _ = R.max(dim=1, keepdim=True)
        logic = (R == maxv).float() if self.is_test else (R > 0.01).float()
        Bm = R * logic
        score_pad = proposal_score.new_zeros(m)
        score_pad[:P] = proposal_score
        ms = (R.clamp(0, 1) * sim_pad).max(1)[0]
        ds = (score_pad.view(1, -1) * Bm).sum(1)
        MS = ops.pad_cols(P, O)
        pad = lambda t: torch.nn.functional.pad(t, (0, MS - m))[None].contiguous()
        return pad(R), pad(Bm), ms[None], ds[None], pad(logic)

    def _forward_hungarian(self, proposed_feature, proposed_mask, template_feature, mask_last_occurence, proposal_score, targets):
        sim, n_prop, n_tplt, match_loss = self.compute_cost_matrix(
            {'proposed': proposed_feature, 'template': template_feature},
            {'proposed': proposed_mask, 'template': mask_last_occurence}, {'proposal_score': proposal_score}, targets)
        full_outmask, ms, ds, _, _ = self.match_with_first_frame(sim, n_prop, n_tplt, proposed_mask.float(), proposal_score,
                                                                 mask_last_occurence)
        return full_outmask, ms, ds, full_outmask, match_loss
