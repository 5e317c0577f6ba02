#!/usr/bin/env python
"""Synthetic train.py-shaped loop (BASELINE.json configs[4]; reference train.py:62-68,186 + dmm/modules/trainer.py:93-300).

What is stock torch here, as the north-star leaves it: a torchvision ResNet-50 (random init) with a 128-channel neck at
strides 4/8/16/32 (reference dmm/modules/base.py:35-54), a small conv decoder standing in for the ConvLSTM refiner, Adam,
and DistributedDataParallel over NCCL for the gradient all-reduce of encoder + neck + decoder (the matching layer has no
parameters, so it adds nothing to the exchange; the reference's second `average_gradients` pass is not repeated).

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
from dmm_net_b200 import ops                                         # noqa: E402
from dmm_net_b200.modules.dmm_model import DMM_Model                # noqa: E402
from dmm_net_b200.synth import default_cfg                          # noqa: E402
from dmm_net_b200.utils.boxlist import BoxList                      # noqa: E402
from dmm_net_b200.utils.masker import Masker                        # noqa: E402


class Encoder(nn.Module):
    """ResNet-50 trunk + 1x1 neck convs to 128 channels per level (base.py:35-54)."""

    def __init__(self, arch="resnet50"):
        super().__init__()
        import torchvision
        net = getattr(torchvision.models, arch)(weights=None)
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        self.layers = nn.ModuleList([net.layer1, net.layer2, net.layer3, net.layer4])
        chans = [256, 512, 1024, 2048] if arch != "resnet18" else [64, 128, 256, 512]
        self.neck = nn.ModuleList([nn.Conv2d(c, 128, 1) for c in chans])

    def forward(self, x):
        x = self.stem(x)
        outs = []
        for layer, neck in zip(self.layers, self.neck):
            x = layer(x)
            outs.append(neck(x))
        return tuple(outs)                                           # strides 4, 8, 16, 32


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


def run(args):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ddp = world > 1
    if ddp:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)                                          # same initial weights on every rank
    H, W = args.size
    B, T, Fo, P = args.clips, args.frames, args.objects, args.proposals
    enc, dec = Encoder(args.arch).to(dev), TinyDecoder().to(dev)
    params = list(enc.parameters()) + list(dec.parameters())
    if ddp:
        from torch.nn.parallel import DistributedDataParallel as DDP
        # several forwards per backward (one per frame / object): BN buffers must not be re-broadcast in between
        enc, dec = DDP(enc, device_ids=[local], broadcast_buffers=False), DDP(dec, device_ids=[local], broadcast_buffers=False)
    opt = torch.optim.Adam(params, lr=1e-4)
    dmm = DMM_Model(default_cfg(10, 5), is_test=0).to(dev)           # train.yaml: 10 x 5 iterations
    masker = Masker(threshold=0.5, padding=1)
    gen = torch.Generator(device=dev).manual_seed(7000 + rank)
    valid = torch.ones(B, Fo, device=dev)
    step_ms, hist = [], []
    for step in range(args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss_total, hard_iou = 0.0, []
        tplt, prev_mask = None, None
        for t in range(T):
            img, gt, prop, m28 = synth_batch(gen, B, Fo, H, W, P, dev)
            feats = enc(img)
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
            outs = []
            for o in range(Fo):                                                                     # the refiner runs per object
                logit = dec(feats, [lv[o] for lv in levels])
                outs.append(F.interpolate(logit, size=(H, W), mode="bilinear", align_corners=False))
            out_masks = torch.sigmoid(torch.cat(outs, 1))                                           # [B,Fo,H,W]
            inter = (out_masks * y_mask).sum((2, 3))
            soft_iou = 1 - inter / ((out_masks + y_mask - out_masks * y_mask).sum((2, 3)) + 1e-6)
            loss_total = loss_total + soft_iou.mean() + sum(match_loss) / len(match_loss)
            with torch.no_grad():
                hard_iou.append(ops.hard_iou_mean(y_mask.flatten(2), out_masks.flatten(2), valid))  # trainer.py:296-300
            prev_mask = out_masks
        loss_total.backward()                                        # DDP all-reduces encoder / neck / decoder grads (NCCL)
        opt.step()
        torch.cuda.synchronize()
        step_ms.append(1e3 * (time.perf_counter() - t0))
        hist.append((float(loss_total.detach()), float(torch.stack(hard_iou).mean())))
        assert hist[-1][0] == hist[-1][0], "loss is NaN"
    g = [p.grad for p in params if p.grad is not None]
    assert len(g) > 0 and all(torch.isfinite(x).all() for x in g)
    if rank == 0:
        ms = sorted(step_ms[1:])[len(step_ms[1:]) // 2] if len(step_ms) > 1 else step_ms[0]
        print(f"{args.arch} ranks={world} {B} clips x {T} frames {H}x{W} P={P} F={Fo}: median step {ms:.1f} ms; "
              f"loss {hist[0][0]:.4f} -> {hist[-1][0]:.4f}; hard IoU {hist[-1][1]:.4f}; {len(g)} parameter tensors with gradients")
    if ddp:
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
    run(ap.parse_args())
