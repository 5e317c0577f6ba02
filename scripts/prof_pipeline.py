"""Eval pipeline per frame, inputs pre-generated: paste-all (Masker -> filter_results -> DMM_Model.inference) against the lazy
pipeline (DMM_Model.inference_lazy).  CUDA events; same outputs (asserted)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200.modules.dmm_model import DMM_Model
from dmm_net_b200.synth import default_cfg
from dmm_net_b200.utils.boxlist import BoxList
from dmm_net_b200.utils.boxlist_ops import filter_results
from dmm_net_b200.utils.masker import Masker

dev = "cuda"
H, W, C, F, P, n_det = 256, 448, 128, 5, 50, 64
for B in (8, 32, 64):
    gen = torch.Generator(device=dev).manual_seed(B)
    model = DMM_Model(default_cfg(40, 5), is_test=1).to(dev)
    valid = torch.ones(B, F, device=dev)
    feats = tuple(torch.randn(B, C, H // s, W // s, generator=gen, device=dev) for s in (4, 8, 16, 32))

    def boxes(n):
        xy = torch.rand(n, 2, generator=gen, device=dev) * torch.tensor([W * 0.6, H * 0.6], device=dev)
        wh = torch.rand(n, 2, generator=gen, device=dev) * torch.tensor([W * 0.3, H * 0.3], device=dev) + W / 8
        return torch.cat([xy, torch.minimum(xy + wh, torch.tensor([W - 1.0, H - 1.0], device=dev))], 1)

    tb = [boxes(F) for _ in range(B)]
    tplt = model.fill_template_dict(None, [BoxList(b) for b in tb], {"backbone_feature": feats, "refine_input_feat": feats}, None, valid)
    masker = Masker(0.5, 1)
    m0, _ = masker([torch.ones(F, 1, 28, 28, device=dev)] * B, [BoxList(b, (W, H)) for b in tb])
    last = torch.stack([m.squeeze(1) for m in m0], 0)
    infos = {"args": None, "shape": (H, W), "extra_frame": [0] * B, "valid": valid}
    dets = []
    for b in range(B):
        d = BoxList(boxes(n_det), (W, H))
        d.add_field("mask", torch.sigmoid(3 * torch.randn(n_det, 1, 28, 28, generator=gen, device=dev) + 1))
        d.add_field("scores", torch.rand(n_det, generator=gen, device=dev))
        dets.append(d)

    def paste_all():
        pasted, tight = masker([d.get_field("mask") for d in dets], dets)
        props = []
        for b, d in enumerate(dets):
            bl = BoxList(tight[b].float(), (W, H))
            bl.add_field("mask", pasted[b])
            bl.add_field("scores", d.get_field("scores"))
            props.append(bl)
        props = filter_results(props, nms_thresh=0.8, max_proposals=P)
        return model.inference(infos, props, feats, last, tplt)[0]

    def lazy():
        return model.inference_lazy(infos, dets, feats, last, tplt, 0.8, P)[0]

    res = {}
    with torch.no_grad():
        for name, fn in (("paste-all", paste_all), ("lazy", lazy)):
            for _ in range(2):
                out = fn()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                out = fn()
            b_.record()
            torch.cuda.synchronize()
            res[name] = (a.elapsed_time(b_) / 5, out)
    assert torch.equal(res["paste-all"][1], res["lazy"][1])
    mb_all = B * (n_det * H * W * 4 * 2 + P * H * W * 4 + F * H * W * 4 * 3) / 1e6     # write+gather pasted, IoU read, templates, out
    mb_lazy = B * (n_det * H * W / 8 + F * H * W * 4 * 2 + P * H * W / 8) / 1e6
    print(f"B={B} clips x {n_det} detections -> {P} proposals, {F} objects, {H}x{W}: paste-all {res['paste-all'][0]:.2f} ms/frame "
          f"(~{mb_all:.0f} MB of mask traffic), lazy {res['lazy'][0]:.2f} ms/frame (~{mb_lazy:.0f} MB); identical outputs", flush=True)
