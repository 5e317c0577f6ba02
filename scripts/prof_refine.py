"""Timings of the rows around the layer (K6 pyramid, K7 labels, K8 paste) with CUDA events; prints GB/s of algorithmic bytes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    H, W = 256, 448
    for B, O in ((4, 5), (32, 5), (128, 10)):
        prev, ref, init = (torch.rand(B, O, H, W, device="cuda") for _ in range(3))
        ms = timeit(lambda: ops.mask_pyramid(prev, ref, init, 4))
        by = 3 * B * O * H * W * 4 * (1 + 1 / 16 + 1 / 64 + 1 / 256 + 1 / 1024)
        print(f"K6 pyramid fwd B={B} O={O}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s")
        init.requires_grad_(True)
        lv = ops.mask_pyramid(prev, ref, init, 4)
        gs = [torch.rand_like(l) for l in lv]
        ms = timeit(lambda: torch.autograd.grad(lv, init, gs, retain_graph=True))
        print(f"K6 pyramid bwd (init only) B={B} O={O}: {ms*1e3:.1f} us  {2*B*O*H*W*4/ms/1e6:.0f} GB/s")
        init.requires_grad_(False)
        pool = torch.nn.MaxPool2d((2, 2), ceil_mode=True)

        def torch_ref():
            for t in range(O):
                m = torch.cat([prev[:, t].reshape(B, 1, H * W), ref[:, t].reshape(B, 1, H * W), init[:, t].reshape(B, 1, H * W)], 2).view(B, 3, H, W)
                m = pool(m)
                for _ in range(4):
                    m = pool(m)
        print(f"   torch op sequence on the same GPU: {timeit(torch_ref)*1e3:.1f} us")
        nv = torch.full((B,), O, device="cuda", dtype=torch.int32)
        outs = torch.rand(B, O, H * W, device="cuda")
        ms = timeit(lambda: ops.merge_labels(outs, nv))
        print(f"K7 labels B={B} O={O}: {ms*1e3:.1f} us  {(B*O*H*W*4+B*H*W)/ms/1e6:.0f} GB/s")
    if hasattr(ops, "paste_masks"):
        for N in (50, 400, 3200):
            masks = torch.rand(N, 1, 28, 28, device="cuda")
            cx, cy = torch.rand(N, device="cuda") * W, torch.rand(N, device="cuda") * H
            bw, bh = torch.rand(N, device="cuda") * W * 0.5 + 8, torch.rand(N, device="cuda") * H * 0.5 + 8
            boxes = torch.stack([(cx - bw / 2).clamp(0, W - 1), (cy - bh / 2).clamp(0, H - 1), (cx + bw / 2).clamp(0, W - 1), (cy + bh / 2).clamp(0, H - 1)], 1)
            ms = timeit(lambda: ops.paste_masks(masks, boxes, H, W))
            print(f"K8 paste N={N}: {ms*1e3:.1f} us  {N*H*W*4/ms/1e6:.0f} GB/s written")
            ms = timeit(lambda: ops.paste_masks(masks, boxes, H, W, want_bits=True))
            print(f"K8 paste+bits N={N}: {ms*1e3:.1f} us")
            ms = timeit(lambda: ops.paste_masks(masks, boxes, H, W, want_pasted=False, want_bits=True))
            print(f"K8 bits + tight boxes only N={N}: {ms*1e3:.1f} us")


if __name__ == "__main__":
    main()
