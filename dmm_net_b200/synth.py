"""Seeded synthetic inputs for the matching layer.

Proposal masks imitate what ``paste_mask_in_image`` produces (reference
``dmm/utils/masker.py:120-155``): soft probabilities inside a tight box, exact zeros outside.
Template masks are either a jittered copy of a proposal's box (so IoU is non-trivial) or a fresh
random blob.  Template features are a chosen proposal's feature plus noise, so the cosine part has
a planted assignment with margin.  Everything is generated on the device of ``device`` from a
``torch.Generator`` seeded with ``1000*config + problem index`` (batched calls use one seed for
the whole batch and are only used for benchmarking).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class MatchProblem:
    prop_feat: torch.Tensor     # [B,P,D] or [P,D]
    prop_mask: torch.Tensor     # [B,P,H,W] or [P,H,W]
    tmpl_feat: torch.Tensor     # [B,O,D] or [O,D]
    tmpl_mask: torch.Tensor     # [B,O,H,W] or [O,H,W]
    prop_score: torch.Tensor    # [B,P] or [P]
    targets: Optional[torch.Tensor] = None
    planted: Optional[torch.Tensor] = None  # [B,O] proposal index each template was derived from

    def to(self, device):
        f = lambda t: None if t is None else t.to(device)
        return MatchProblem(f(self.prop_feat), f(self.prop_mask), f(self.tmpl_feat), f(self.tmpl_mask),
                            f(self.prop_score), f(self.targets), f(self.planted))

    def squeeze0(self):
        f = lambda t: None if t is None else t[0]
        return MatchProblem(f(self.prop_feat), f(self.prop_mask), f(self.tmpl_feat), f(self.tmpl_mask),
                            f(self.prop_score), f(self.targets), f(self.planted))


def _blobs(boxes: torch.Tensor, H: int, W: int, gen: torch.Generator, soft: bool = True) -> torch.Tensor:
    """boxes [..., 4] = (y0, x0, h, w) float -> masks [..., H, W]: soft blob inside the box, 0 outside."""
    dev = boxes.device
    yy = torch.arange(H, device=dev, dtype=torch.float32).view(*([1] * (boxes.dim() - 1)), H, 1)
    xx = torch.arange(W, device=dev, dtype=torch.float32).view(*([1] * (boxes.dim() - 1)), 1, W)
    y0, x0, h, w = [boxes[..., i, None, None] for i in range(4)]
    inside = (yy >= y0) & (yy < y0 + h) & (xx >= x0) & (xx < x0 + w)
    ry = (yy - (y0 + 0.5 * h)) / (0.5 * h)
    rx = (xx - (x0 + 0.5 * w)) / (0.5 * w)
    val = (1.2 - (ry * ry + rx * rx)).clamp_(0.0, 1.0)
    if soft:
        noise = torch.empty(val.shape, device=dev, dtype=torch.float32).uniform_(0.6, 1.0, generator=gen)
        val = val * noise
    return val * inside


def make_problems(B: int, P: int, O: int, H: int, W: int, D: int = 512, seed: int = 0, device="cpu",
                  with_targets: bool = False, dup_frac: float = 0.0, chunk: int = 16) -> MatchProblem:
    """A batch of independent (proposals, templates) problems, all of the same shape."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    U = lambda *s: torch.rand(*s, device=dev, generator=gen)

    ph = (H / 8 + U(B, P) * (H / 2 - H / 8)).floor().clamp_(min=1)
    pw = (W / 8 + U(B, P) * (W / 2 - W / 8)).floor().clamp_(min=1)
    py = (U(B, P) * (H - ph)).floor()
    px = (U(B, P) * (W - pw)).floor()
    pbox = torch.stack([py, px, ph, pw], -1)                              # [B,P,4]

    planted = torch.stack([torch.randperm(P, device=dev, generator=gen)[:O] if P >= O else
                           torch.randint(0, P, (O,), device=dev, generator=gen) for _ in range(B)], 0)  # [B,O]
    src = torch.gather(pbox, 1, planted[..., None].expand(-1, -1, 4))     # [B,O,4]
    jit = (U(B, O, 4) - 0.5) * torch.tensor([H / 16, W / 16, H / 16, W / 16], device=dev)
    tbox_a = src + jit.round()
    th = (H / 8 + U(B, O) * (H / 2 - H / 8)).floor().clamp_(min=1)
    tw = (W / 8 + U(B, O) * (W / 2 - W / 8)).floor().clamp_(min=1)
    tbox_b = torch.stack([(U(B, O) * (H - th)).floor(), (U(B, O) * (W - tw)).floor(), th, tw], -1)
    use_a = (U(B, O) < 0.5)[..., None]
    tbox = torch.where(use_a, tbox_a, tbox_b)
    tbox[..., 2:] = tbox[..., 2:].clamp(min=1)

    prop_mask = torch.empty(B, P, H, W, device=dev)
    tmpl_mask = torch.empty(B, O, H, W, device=dev)
    tgt = torch.empty(B, O, H, W, device=dev) if with_targets else None
    for s in range(0, B, chunk):                                          # bound the temporaries
        e = min(B, s + chunk)
        prop_mask[s:e] = _blobs(pbox[s:e], H, W, gen)
        tmpl_mask[s:e] = _blobs(tbox[s:e], H, W, gen)
        if with_targets:
            tgt[s:e] = (_blobs(src[s:e], H, W, gen, soft=False) > 0.3).float()

    prop_feat = torch.randn(B, P, D, device=dev, generator=gen)
    tmpl_feat = torch.gather(prop_feat, 1, planted[..., None].expand(-1, -1, D)) \
        + 0.3 * torch.randn(B, O, D, device=dev, generator=gen)
    score = U(B, P)
    if dup_frac > 0:                                                      # exact duplicate proposals -> ties
        ndup = max(1, int(P * dup_frac))
        prop_mask[:, P - ndup:] = prop_mask[:, :ndup]
        prop_feat[:, P - ndup:] = prop_feat[:, :ndup]
    return MatchProblem(prop_feat, prop_mask, tmpl_feat, tmpl_mask, score, tgt, planted)


def make_problem(P: int, O: int, H: int, W: int, D: int = 512, config: int = 0, index: int = 0, device="cpu",
                 with_targets: bool = False, dup_frac: float = 0.0) -> MatchProblem:
    """One problem with the section-8(d) seed rule: seed = 1000*config + index."""
    return make_problems(1, P, O, H, W, D, 1000 * config + index, device, with_targets, dup_frac).squeeze0()


def default_cfg(max_iter: int = 20, proj_iter: int = 5, lr: float = 0.1, score_weight: float = 0.3, algo: str = "relax"):
    """The five keys MatchModel reads (reference match_model.py:18-21,90; presets in dmm/configs/*.yaml)."""
    return {"matching": {"algo": algo, "match_max_score": 1, "cost": "cosine"}, "relax_max_iter": max_iter,
            "relax_proj_iter": proj_iter, "relax_learning_rate": lr, "score_weight": score_weight}
