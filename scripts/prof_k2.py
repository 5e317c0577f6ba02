"""K2 cosine: tensor-core (DMM_K2_IMPL=tc) vs SIMT (DMM_K2_IMPL=simt) -- accuracy against an fp64 reference and CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200 import ops


def ref64(q, k, eps=1e-8):
    q, k = q.double(), k.double()
    qn = q.norm(dim=-1).clamp_min(eps)[..., :, None]
    kn = k.norm(dim=-1).clamp_min(eps)[..., None, :]
    return (q @ k.transpose(-1, -2)) / (qn * kn)


def main():
    impl = os.environ.get("DMM_K2_IMPL", "default")
    torch.manual_seed(0)
    cases = ((1, 50, 10, 512), (2, 50, 10, 512), (7, 13, 5, 64), (64, 50, 10, 512), (1024, 50, 10, 512), (333, 64, 16, 260), (5, 1, 1, 32), (8192, 50, 10, 512))
    if "--big" in sys.argv:
        cases = ((1024, 50, 10, 512), (8192, 50, 10, 512))
    for B, P, O, D in cases:
        k = torch.randn(B, P, D, device="cuda")
        q = k[:, torch.arange(O) % P] + 0.3 * torch.randn(B, O, D, device="cuda")
        if B > 2:
            k[1, 0].zero_()                                  # zero vector -> eps clamp
            k[2] *= 1e3
        cos = ops.cosine_pairwise(q[:, None], k)
        torch.cuda.synchronize()
        err = (cos.double() - ref64(q, k)).abs().max().item()
        for _ in range(3):
            ops.cosine_pairwise(q[:, None], k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 20
        for _ in range(n):
            ops.cosine_pairwise(q[:, None], k)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print(f"[{impl}] B={B} P={P} O={O} D={D}: max|err| vs fp64 {err:.2e}   {ms*1e3:.1f} us   {B*(P+O)*D*4/ms/1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
