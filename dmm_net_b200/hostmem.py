"""Host-memory placement for the host-buffer entry (ops.match_batch_host): on a two-socket box every GPU hangs off one
socket; pinned buffers allocated on the OTHER socket make both routes (packing threads and copy-engine DMA) cross the
socket interconnect, and buffers of all ranks allocated on ONE socket share that socket's DRAM channels.  These helpers
find the GPU's NUMA node and allocate pinned tensors with a `preferred node` memory policy (set_mempolicy through libc's
syscall(); no libnuma in the image).  Everything degrades to plain allocation when the kernel / cgroup says no."""
from __future__ import annotations

import ctypes
import os
from contextlib import contextmanager
from typing import Optional

import torch

_SYS_SET_MEMPOLICY = 238          # x86_64
_MPOL_DEFAULT, _MPOL_PREFERRED = 0, 1


def gpu_numa_node(index: int) -> Optional[int]:
    """NUMA node of CUDA device `index` from sysfs (None when unknown / single node / -1)."""
    bdf = None
    try:
        pr = torch.cuda.get_device_properties(index)
        if hasattr(pr, "pci_bus_id") and hasattr(pr, "pci_device_id"):
            bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    except Exception:
        bdf = None
    if bdf is None:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(_physical_index(index))
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        except Exception:
            return None
    bdf = str(bdf).lower()
    if len(bdf.split(":")[0]) == 8:                                  # nvml prints an 8-digit domain, sysfs uses 4
        bdf = bdf[4:]
    try:
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
    except Exception:
        return None
    return node if node >= 0 else None


def _physical_index(index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[index])
        except Exception:
            pass
    return index


def memory_nodes() -> list:
    """NUMA nodes this process may allocate from (cgroup cpuset.mems.effective when present, else all online nodes)."""
    def parse(txt):
        out = []
        for part in txt.strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            out.extend(range(int(a), int(b or a) + 1))
        return out
    for path in ("/sys/fs/cgroup/cpuset.mems.effective", "/sys/devices/system/node/online"):
        try:
            nodes = parse(open(path).read())
            if nodes:
                return nodes
        except Exception:
            continue
    return [0]


@contextmanager
def prefer_node(node: Optional[int]):
    """Allocations whose pages are first touched inside the block go to `node` when the kernel allows it."""
    ok = False
    if node is not None and node in memory_nodes():
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            ok = libc.syscall(_SYS_SET_MEMPOLICY, _MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask))) == 0
        except Exception:
            ok = False
    try:
        yield ok
    finally:
        if ok:
            try:
                ctypes.CDLL(None).syscall(_SYS_SET_MEMPOLICY, _MPOL_DEFAULT, None, ctypes.c_ulong(0))
            except Exception:
                pass


def pinned_near_gpu(t: torch.Tensor, device_index: int):
    """A pinned copy of CPU tensor `t` whose pages sit on the GPU's NUMA node when that can be arranged.
    Returns (tensor, node or None)."""
    node = gpu_numa_node(device_index)
    with prefer_node(node) as ok:
        out = torch.empty(t.shape, dtype=t.dtype).pin_memory()        # cudaHostAlloc + first touch under the policy
        out.copy_(t)
    return out, (node if ok else None)
