"""Host packer throughput vs thread count on this box (decides whether packing before PCIe helps e2e)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops
m = torch.rand(8 * 60, 256 * 448).pin_memory()          # 8 matches worth of masks, 220 MB
out = torch.empty(8 * 60, 3584, dtype=torch.int32).pin_memory()
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for th in (1, 2, 4, 8, 16, 32, 64, 128, 0):
    ops.pack_masks_host(m, mask_dims=1, threads=th, out=out)
    t = time.perf_counter()
    n = 3
    for _ in range(n):
        ops.pack_masks_host(m, mask_dims=1, threads=th, out=out)
    dt = (time.perf_counter() - t) / n
    print(f"threads={th:4d}: {dt * 1e3:8.2f} ms  {m.numel() * 4 / dt / 1e9:7.1f} GB/s  -> {8 / dt:8.0f} matches/s")
