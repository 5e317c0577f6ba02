#!/usr/bin/env python
"""Synthetic train.py-shaped loop (BASELINE.json configs[4]; reference train.py:62-68,186 + dmm/modules/trainer.py:93-300).

What is stock torch here, as the north-star leaves it: a torchvision ResNet-50 (random init) with a 128-channel neck at
strides 4/8/16/32 (reference dmm/modules/base.py:35-54), a small conv decoder standing in for the ConvLSTM refiner, Adam,
and the data-parallel gradient exchange of encoder + neck + decoder over NCCL (the matching layer has no parameters, so
it adds nothing to the exchange).  Two exchanges are implemented: ``--reduce ddp`` = torch DistributedDataParallel as in
train.py:178-184 (bucketed all-reduce overlapped with backward), ``--reduce flat`` (default for N > 1) = every gradient is
a view into ONE flat buffer that a single NCCL all-reduce averages after backward -- no per-parameter hooks or messages;
the reference's second per-parameter `average_gradients` pass (train.py:62-68) is not repeated in either.

What runs through this repo's kernels, with autograd: K8 proposal paste (no grad), K5 ROI mean pooling (grad into the
feature maps), the batched DMM_Model container in training mode -- K2 cosine (tcgen05 forward, fp32 backward), K1
mask-IoU against the previous masks and against the targets in one pass, K3 relaxed solver with its backward, K4
assignment apply with its backward -- K6 decoder mask-input pyramid for every object at once (grad into the matched
masks), the hard-IoU metric (K1 row-wise).  Clips are sharded over ranks; B clips x T frames per step per rank
(scripts/train/train_r50.sh: 4 x 3).

  python examples/synthetic_train_step.py [--steps 5] [--clips 4] [--frames 3] [--size 256 448]
  torchrun --nproc-per-node N --master-addr 127.0.0.1 examples/synthetic_train_step.py ...
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from backbone import Encoder                                        # noqa: E402
from dmm_net_b200 import ops                                         # noqa: E402
from dmm_net_b200.modules.dmm_model import DMM_Model                # noqa: E402
from dmm_net_b200.sharding import FlatGradBucket                   # noqa: E402
from dmm_net_b200.synth import default_cfg                          # noqa: E402
from dmm_net_b200.utils.boxlist import BoxList                      # noqa: E402
from dmm_net_b200.utils.masker import Masker                        # noqa: E402


class TinyDecoder(nn.Module):
    """Stand-in for the ConvLSTM refiner (trainer.py:236-300): per object, fuses each feature level with the 3-channel
    mask input of that level (K6 output) coarse-to-fine and predicts a full-resolution mask logit."""

    def __init__(self):
        super().__init__()
        self.fuse = nn.ModuleList([nn.Conv2d(128 + 3 + (16 if i else 0), 16, 3, padding=1) for i in range(4)])
        self.head = nn.Conv2d(16, 1, 1)

    def forward(self, feats, mask_levels):
        """feats: 4 x [B,128,h,w] fine->coarse; mask_levels: 4 x [B,3,h,w] fine->coarse"""
        h = None
        for i, lvl in enumerate(reversed(range(4))):                 # coarse -> fine, like reversed(mask_lstm)
            x = torch.cat([feats[lvl], mask_levels[lvl]] + ([F.interpolate(h, size=feats[lvl].shape[-2:])] if h is not None else []), 1)
            h = torch.relu(self.fuse[i](x))
        return self.head(h)                                          # [B,1,H/4,W/4]


def synth_batch(gen, B, Fo, H, W, P, dev):
    """images, ground-truth object boxes / masks, and per-frame proposals as the mask head would deliver them"""
    img = torch.randn(B, 3, H, W, generator=gen, device=dev)
    x1 = torch.rand(B, Fo, generator=gen, device=dev) * W * 0.6
    y1 = torch.rand(B, Fo, generator=gen, device=dev) * H * 0.6
    gt = torch.stack([x1, y1, (x1 + W * 0.3).clamp(max=W - 1), (y1 + H * 0.3).clamp(max=H - 1)], 2)   # [B,Fo,4]
    jit = torch.randn(B, P, 4, generator=gen, device=dev) * 6
    prop = gt[:, torch.arange(P, device=dev) % Fo] + jit                                               # proposals around the objects
    prop[..., 0::2] = prop[..., 0::2].clamp(0, W - 1)
    prop[..., 1::2] = prop[..., 1::2].clamp(0, H - 1)
    prop = torch.cat([torch.minimum(prop[..., :2], prop[..., 2:] - 2), prop[..., 2:]], -1).clamp(min=0)
    m28 = torch.sigmoid(4 * torch.randn(B, P, 1, 28, 28, generator=gen, device=dev) + 2)
    return img, gt, prop, m28


def train_loop(arch="resnet50", clips=4, frames=3, objects=3, proposals=50, size=(256, 448), steps=5, warmup=1,
               reduce="auto", fused_adam=True, seed_offset=0):
    """Runs warmup + steps training steps; returns dict(hist, step_ms (CUDA events, per step), host_ms, fwd_ms, bwd_ms,
    reduce_ms, opt_ms, grad_bytes, n_grad_tensors, reduce).

    reduce: "none" (no exchange), "ddp" (torch DistributedDataParallel: bucketed all-reduce overlapped with backward),
    "flat" (FlatGradBucket: one all-reduce after backward), "auto" = "flat" when world > 1 else "none"."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.device("cuda", torch.cuda.current_device())
    if reduce == "auto":
        reduce = "flat" if world > 1 else "none"
    if world == 1 and reduce == "ddp":
        reduce = "none"
    torch.manual_seed(1234)                                          # same initial weights on every rank
    H, W = size
    B, T, Fo, P = clips, frames, objects, proposals
    enc, dec = Encoder(arch).to(dev), TinyDecoder().to(dev)
    params = list(enc.parameters()) + list(dec.parameters())
    enc_f, dec_f = enc, dec
    bucket = None
    if reduce == "ddp":
        from torch.nn.parallel import DistributedDataParallel as DDP
        # several forwards per backward (one per frame / object): BN buffers must not be re-broadcast in between; the
        # gradients live inside the buckets (no copy in, no copy out)
        kw = dict(device_ids=[dev.index], broadcast_buffers=False, gradient_as_bucket_view=True, bucket_cap_mb=64)
        enc_f, dec_f = DDP(enc, **kw), DDP(dec, **kw)
    elif reduce == "flat" or reduce == "none":
        bucket = FlatGradBucket(params)
    opt = torch.optim.Adam(params, lr=1e-4, fused=bool(fused_adam))
    dmm = DMM_Model(default_cfg(10, 5), is_test=0).to(dev)           # train.yaml: 10 x 5 iterations
    masker = Masker(threshold=0.5, padding=1)
    gen = torch.Generator(device=dev).manual_seed(7000 + rank + seed_offset)
    valid = torch.ones(B, Fo, device=dev)
    hist = []
    rec = {k: [] for k in ("step_ms", "host_ms", "fwd_ms", "bwd_ms", "skew_ms", "reduce_ms", "opt_ms")}
    E = lambda: torch.cuda.Event(enable_timing=True)
    pending = []
    for step in range(warmup + steps):
        e0, e1, e2, e2b, e3, e4 = E(), E(), E(), E(), E(), E()
        t0 = time.perf_counter()
        e0.record()
        if bucket is not None:
            bucket.zero()
        else:
            opt.zero_grad(set_to_none=False)
        loss_total, hard_iou = 0.0, []
        tplt, prev_mask = None, None
        for t in range(T):
            img, gt, prop, m28 = synth_batch(gen, B, Fo, H, W, P, dev)
            feats = enc_f(img)
            gt_lists = [BoxList(gt[b], (W, H)) for b in range(B)]
            gt_masks, _ = masker([torch.ones(Fo, 1, 28, 28, device=dev)] * B, gt_lists)           # K8: ground-truth masks
            y_mask = torch.stack([m.squeeze(1) for m in gt_masks], 0)                               # [B,Fo,H,W]
            if t == 0:
                tplt = dmm.fill_template_dict(None, gt_lists, {"backbone_feature": feats, "refine_input_feat": feats}, None, valid)
                prev_mask = y_mask
                ref_mask = y_mask
                continue
            props = []
            lists = [BoxList(prop[b], (W, H)) for b in range(B)]
            pasted, tight = masker([m28[b] for b in range(B)], lists)                               # K8: every proposal, one launch
            for b in range(B):
                bl = BoxList(tight[b].float(), (W, H))
                bl.add_field("mask", pasted[b])
                bl.add_field("scores", torch.rand(P, generator=gen, device=dev))
                props.append(bl)
            init_pred, tplt, match_loss, _ = dmm(None, props, feats, prev_mask.detach(), tplt, valid, y_mask)
            levels = ops.mask_pyramid(prev_mask.detach(), ref_mask, init_pred, 4)                   # K6: all objects at once
            # the refiner runs per object in the reference (trainer.py:236-300); the objects are independent given the
            # frame's features, so they go through the decoder as ONE batch of Fo*B maps (one forward per frame)
            logit = dec_f([f.repeat(Fo, 1, 1, 1) for f in feats], [lv.flatten(0, 1) for lv in levels])   # [Fo*B,1,h,w]
            logit = F.interpolate(logit, size=(H, W), mode="bilinear", align_corners=False)
            out_masks = torch.sigmoid(logit.view(Fo, B, H, W).transpose(0, 1))                      # [B,Fo,H,W]
            inter = (out_masks * y_mask).sum((2, 3))
            soft_iou = 1 - inter / ((out_masks + y_mask - out_masks * y_mask).sum((2, 3)) + 1e-6)
            loss_total = loss_total + soft_iou.mean() + sum(match_loss) / len(match_loss)
            with torch.no_grad():
                hard_iou.append(ops.hard_iou_mean(y_mask.flatten(2), out_masks.flatten(2), valid))  # trainer.py:296-300
            prev_mask = out_masks
        e1.record()
        loss_total.backward()                                        # "ddp": bucketed NCCL all-reduce overlapped with this
        e2.record()
        if reduce == "flat":
            bucket.rendezvous()                                      # wait for the slowest rank (timed separately)
        e2b.record()
        if reduce == "flat":
            bucket.all_reduce_mean()                                 # ONE all-reduce of every gradient (NCCL, NVLink)
        e3.record()
        opt.step()
        e4.record()
        host_ms = 1e3 * (time.perf_counter() - t0)
        pending.append((step, e0, e1, e2, e2b, e3, e4, host_ms, loss_total.detach(), torch.stack(hard_iou).mean()))
    torch.cuda.synchronize()
    for step, e0, e1, e2, e2b, e3, e4, host_ms, loss, hi in pending:
        hist.append((float(loss), float(hi)))
        assert hist[-1][0] == hist[-1][0], "loss is NaN"
        if step < warmup:
            continue
        rec["step_ms"].append(e0.elapsed_time(e4))
        rec["host_ms"].append(host_ms)
        rec["fwd_ms"].append(e0.elapsed_time(e1))
        rec["bwd_ms"].append(e1.elapsed_time(e2))
        rec["skew_ms"].append(e2.elapsed_time(e2b))
        rec["reduce_ms"].append(e2b.elapsed_time(e3))
        rec["opt_ms"].append(e3.elapsed_time(e4))
    g = [p.grad for p in params if p.grad is not None]
    assert len(g) > 0 and all(torch.isfinite(x).all() for x in g)
    med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")
    out = {k: med(v) for k, v in rec.items()}
    out.update(hist=hist, reduce=reduce, n_grad_tensors=len(g), grad_bytes=4 * sum(x.numel() for x in g),
               clips=B, frames=T, objects=Fo, proposals=P, size=(H, W), arch=arch, world=world)
    return out


def run(args):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    r = train_loop(args.arch, args.clips, args.frames, args.objects, args.proposals, tuple(args.size), args.steps,
                   getattr(args, "warmup", 0), getattr(args, "reduce", "auto"))
    hist = r["hist"]
    if rank == 0:
        print(f"{args.arch} ranks={world} reduce={r['reduce']} {r['clips']} clips x {r['frames']} frames {args.size[0]}x{args.size[1]} "
              f"P={r['proposals']} F={r['objects']}: median step {r['step_ms']:.1f} ms (host {r['host_ms']:.1f}; fwd {r['fwd_ms']:.1f} "
              f"bwd {r['bwd_ms']:.1f} skew {r['skew_ms']:.2f} reduce {r['reduce_ms']:.2f} opt {r['opt_ms']:.2f}); loss {hist[0][0]:.4f} -> {hist[-1][0]:.4f}; "
              f"hard IoU {hist[-1][1]:.4f}; {r['n_grad_tensors']} parameter tensors with gradients, {r['grad_bytes'] / 1e6:.1f} MB")
    if world > 1:
        dist.destroy_process_group()
    return hist


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--clips", type=int, default=4)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--objects", type=int, default=3)
    ap.add_argument("--proposals", type=int, default=50)
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--size", type=int, nargs=2, default=[256, 448])
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--reduce", default="auto", choices=["auto", "none", "ddp", "flat"])
    run(ap.parse_args())
