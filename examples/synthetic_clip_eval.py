#!/usr/bin/env python
"""Clip-sharded inference loop shaped like the reference's eval path (eval.py -> Evaler.forward ->
DMM_Model.inference, reference dmm/modules/evaluator.py:83-134): frames of a clip are sequential (the matched masks of
frame t are the templates of frame t+1), clips are independent and sharded over ranks (eval.py:57-59).

Everything outside the matching path is synthetic here: the backbone features are random 128-channel maps at strides
4/8/16/32 (the north-star leaves the backbone on stock torch convs), proposals are random boxes with random 28x28 mask
head outputs, and the decoder is the identity.  What runs for real is the scope of this repo, one launch per kernel per
frame for all clips of the rank: K8 pastes every proposal mask into the image (Masker), K9 runs the per-frame NMS on
the tight boxes (filter_results), K5 pools the proposal features, the batched DMM_Model container matches (K2 cosine on
the tensor cores, K1 mask-IoU through the per-video pointer table, K3 solver, K4 apply with the valid-row scatter), K6
builds the decoder's mask-input pyramid for every object and K7 merges the output masks into the label map.

  python examples/synthetic_clip_eval.py [--clips 8] [--frames 12] [--proposals 50] [--objects 5] [--size 256 448]
  torchrun --nproc-per-node N examples/synthetic_clip_eval.py ...      (clips sharded, no collective in the data path)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dmm_net_b200 import ops                                    # noqa: E402
from dmm_net_b200.modules.dmm_model import DMM_Model            # noqa: E402
from dmm_net_b200.utils.boxlist_ops import filter_results      # noqa: E402
from dmm_net_b200.utils.masker import Masker                    # noqa: E402
from dmm_net_b200.sharding import aggregate_throughput, shard_indices   # noqa: E402
from dmm_net_b200.synth import default_cfg                      # noqa: E402
from dmm_net_b200.utils.boxlist import BoxList                  # noqa: E402


def random_boxes(gen, n, H, W, dev):
    x1 = torch.rand(n, generator=gen, device=dev) * (W * 0.7)
    y1 = torch.rand(n, generator=gen, device=dev) * (H * 0.7)
    w = W / 8 + torch.rand(n, generator=gen, device=dev) * (W / 3)
    h = H / 8 + torch.rand(n, generator=gen, device=dev) * (H / 3)
    return torch.stack([x1, y1, (x1 + w).clamp(max=W - 1), (y1 + h).clamp(max=H - 1)], 1)


def paste_masks(boxes, H, W, gen):
    """soft blob inside the box, exact zero outside (what reference masker.py:120-155 produces)"""
    dev = boxes.device
    yy = torch.arange(H, device=dev, dtype=torch.float32).view(1, H, 1)
    xx = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, W)
    x1, y1, x2, y2 = [boxes[:, i].view(-1, 1, 1) for i in range(4)]
    inside = (yy >= y1) & (yy <= y2) & (xx >= x1) & (xx <= x2)
    ry = (yy - (y1 + y2) / 2) / ((y2 - y1) / 2 + 1e-3)
    rx = (xx - (x1 + x2) / 2) / ((x2 - x1) / 2 + 1e-3)
    val = (1.2 - (ry * ry + rx * rx)).clamp(0, 1)
    return (val * inside).unsqueeze(1)                              # [n,1,H,W] like BoxList 'mask'


def clip_eval(clips, frames, proposals=50, objects=5, size=(256, 448), lazy=False, arch=None, fixed_objects=False,
              time_ops=False, seed=4000, graph_backbone=True):
    """Runs `clips` clips of `frames` frames on THIS rank as one batch per frame; frame 0 is the warm-up.
    arch: None (random feature maps) or a torchvision ResNet name (images -> Encoder -> features, stock torch).
    graph_backbone: replay the backbone forward as ONE CUDA graph per frame (stock torch.cuda.CUDAGraph): eager, the
    ~300 small conv / bn / relu launches of a ResNet on 8 images are bound by the host thread issuing them, which is what
    stops the clip loop from scaling over the GPUs of a box with few host cores.
    Returns dict(frames, ms, ms_per_frame_step, out_shape, op_ms {op: total ms} when time_ops)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    H, W = size
    F, C, B = objects, 128, clips
    model = DMM_Model(default_cfg(40, 5), is_test=1).to(dev)         # eval.yaml: 40 x 5 iterations
    gen = torch.Generator(device=dev).manual_seed(seed)
    enc = None
    if arch is not None:
        from backbone import Encoder
        torch.manual_seed(1234)
        enc = Encoder(arch).to(dev).eval()
    n_obj = torch.full((B,), F, device=dev) if fixed_objects else torch.randint(1, F + 1, (B,), generator=gen, device=dev)
    valid = (torch.arange(F, device=dev)[None, :] < n_obj[:, None]).float()

    graph = None
    if enc is not None and graph_backbone:
        img_static = torch.randn(B, 3, H, W, generator=gen, device=dev)
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm), torch.no_grad():             # cuDNN picks its algorithms outside the capture
            for _ in range(3):
                enc(img_static)
        torch.cuda.current_stream().wait_stream(warm)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            feats_static = enc(img_static)

    def feats():
        if enc is None:
            return tuple(torch.randn(B, C, -(-H // s), -(-W // s), generator=gen, device=dev) for s in (4, 8, 16, 32))
        if graph is None:
            return enc(torch.randn(B, 3, H, W, generator=gen, device=dev))
        img_static.normal_(generator=gen)                           # the frame's images land in the graph's input buffer
        graph.replay()
        return feats_static                                         # overwritten by the next replay: consumed within the frame

    timer = ops.KernelTimer() if time_ops else None
    backbone_ev = []
    # frame 0: ground-truth boxes/masks define the templates
    tboxes = [random_boxes(gen, F, H, W, dev) for _ in range(B)]
    with torch.no_grad():
        f0 = tuple(f.clone() for f in feats())                       # templates keep views of frame 0's features
        tplt = model.fill_template_dict(None, [BoxList(b) for b in tboxes], {"backbone_feature": f0, "refine_input_feat": f0},
                                        None, valid)
    mask_last = torch.stack([paste_masks(b, H, W, gen).squeeze(1) for b in tboxes], 0) * valid[:, :, None, None]
    mask0 = mask_last                                                # reference masks of frame 0 (y_mask)
    masker = Masker(threshold=0.5, padding=1)
    infos = {"args": None, "shape": (H, W), "extra_frame": [0] * B, "valid": valid}
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    checks = []
    with torch.no_grad():
        for t in range(frames):
            if t == 1:
                t0.record()                                          # frame 0 is the warm-up
                ops.set_kernel_timer(timer)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fb = feats()
            e1.record()
            if t >= 1:
                backbone_ev.append((e0, e1))
            n_raw = proposals + 14                                   # detections before NMS
            raw = [BoxList(random_boxes(gen, n_raw, H, W, dev), (W, H)) for _ in range(B)]
            m28 = [torch.sigmoid(3 * torch.randn(n_raw, 1, 28, 28, generator=gen, device=dev) + 1.5) for _ in range(B)]
            scores = [torch.rand(n_raw, generator=gen, device=dev) for _ in range(B)]
            prev = mask_last
            if lazy:
                # lazy pipeline: bits-only paste, device-side keep table, packed K1, K10 pastes only the matched detections
                for b in range(B):
                    raw[b].add_field("mask", m28[b])
                    raw[b].add_field("scores", scores[b])
                out, tplt, _, mask_last, _ = model.inference_lazy(infos, raw, fb, mask_last, tplt, 0.8, proposals)
            else:
                pasted, tight = masker(m28, raw)                     # K8: all B x n_raw proposals in one launch
                props = []
                for b in range(B):
                    bl = BoxList(tight[b].float(), (W, H))           # mask post-processor: boxes become the tight boxes
                    bl.add_field("mask", pasted[b])
                    bl.add_field("scores", scores[b])
                    props.append(bl)
                props = filter_results(props, nms_thresh=0.8, max_proposals=proposals)  # K9: one launch per frame
                out, tplt, _, mask_last = model.inference(infos, props, fb, mask_last, tplt)
            levels = ops.mask_pyramid(prev, mask0, out, 4)           # K6: decoder inputs of every object (identity decoder here)
            labels = ops.merge_labels(out.view(B, F, -1), n_obj)     # K7: merged label map (evaluator.py:139-145)
            checks = checks[-1:] + [(out, levels[-1], labels)]   # keep two frames alive, not the clip: a steady allocation pattern
        t1.record()
    ops.set_kernel_timer(None)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) if frames > 1 else float("nan")
    sums = [float(o.sum()) + float(l.sum()) + float(lb.sum()) for o, l, lb in checks[-2:]]
    assert out.shape == (B, F, H, W) and all(c == c for c in sums)
    assert labels.shape == (B, H * W) and int(labels.max()) <= F and levels[0].shape == (F, B, 3, (H + 3) // 4, (W + 3) // 4)
    assert float((out * (1 - valid)[:, :, None, None]).abs().sum()) == 0.0, "rows of invalid templates must stay zero"
    res = {"frames": B * max(frames - 1, 0), "ms": ms, "ms_per_frame_step": ms / max(frames - 1, 1), "checksum": sums[-1]}
    if timer is not None:
        op_ms = timer.totals()
        op_ms["backbone (stock torch)" if enc is not None else "synthetic features"] = sum(a.elapsed_time(b) for a, b in backbone_ev)
        res["op_ms"] = op_ms
    return res


def run(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    mine = shard_indices(args.clips, rank, world)                    # this rank's clips, processed as one batch per frame
    r = clip_eval(len(mine), args.frames, args.proposals, args.objects, tuple(args.size), args.lazy,
                  getattr(args, "arch", None), seed=4000 + rank)
    rate = aggregate_throughput(r["frames"], r["ms"], dev) if args.frames > 1 else 0.0
    if rank == 0:
        print(f"{'lazy' if args.lazy else 'paste-all'} pipeline: clips={args.clips} ranks={world} frames/clip={args.frames} P~{args.proposals} "
              f"F={args.objects} {args.size[0]}x{args.size[1]}: {rate:.0f} (clip,frame) matches/s incl. paste + NMS + pyramid + labels; "
              f"last checksum {r['checksum']:.3f}")
    if world > 1:
        dist.destroy_process_group()
    return rate


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--proposals", type=int, default=50)
    ap.add_argument("--objects", type=int, default=5)
    ap.add_argument("--size", type=int, nargs=2, default=[256, 448])
    ap.add_argument("--lazy", action="store_true", help="never materialise the pasted proposal masks (DMM_Model.inference_lazy)")
    ap.add_argument("--arch", default=None, help="torchvision ResNet for the backbone (default: random feature maps)")
    run(ap.parse_args())
