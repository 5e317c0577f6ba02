"""K5-TC stress under NVLink / HBM traffic: every rank hammers roi_mean_pool (train-loop-like shapes) on its default stream
while a side stream keeps large NCCL all-reduces in flight.  Hunting the raw-ring phase-parity race."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dmm_net_b200 import ops
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
torch.manual_seed(rank)
H, W, C, N = 256, 448, 128, 4
feats = [torch.randn(N, C, H // s, W // s, device=dev) for s in (4, 8, 16, 32)]
big = torch.randn(64 * 1024 * 1024, device=dev)          # 256 MB
side = torch.cuda.Stream()
t0 = time.time()
for it in range(iters):
    if it % 4 == 0:
        with torch.cuda.stream(side):
            dist.all_reduce(big, async_op=True)
    R = 12 if it % 2 == 0 else 200
    per = R // N
    x1 = torch.rand(R, device=dev) * W * 0.6; y1 = torch.rand(R, device=dev) * H * 0.6
    rois = torch.stack([torch.arange(N, device=dev).repeat_interleave(per).float(), x1, y1, x1 + W * 0.3, y1 + H * 0.3], 1)
    out = ops.roi_mean_pool(feats, rois, impl="tc")
    if it % 1000 == 999:
        torch.cuda.synchronize()
        if rank == 0:
            print("iter", it + 1, "ok", round(time.time() - t0, 1), "s", flush=True)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("done", iters, "calls per rank on", world, "ranks")
dist.destroy_process_group()
