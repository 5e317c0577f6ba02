"""Host-buffer entry: sweep of the raw fraction (fp32 over PCIe by the copy engine vs bits packed by the host cores)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmm_net_b200.modules.match_model import MatchModel
from dmm_net_b200.synth import default_cfg, make_problems

B, P, O, H, W, D = 64, 50, 10, 256, 448, 512
pr = make_problems(B, P, O, H, W, D, seed=1, device="cuda")
host = {k: getattr(pr, k).cpu().pin_memory() for k in ("prop_feat", "prop_mask", "tmpl_feat", "tmpl_mask", "prop_score")}
layer = MatchModel(default_cfg(20, 5, 0.1, 0.3), is_test=1)
res = torch.empty(B, O, 50, pin_memory=True)
for th in [int(x) for x in os.environ.get("E2E_THREADS", "8,16,32").split(",")]:
    for f in [None if x == "auto" else float(x) for x in os.environ.get("E2E_FRACS", "0.0,0.2,0.3,0.4,0.5,auto").split(",")]:
        iters = int(os.environ.get("E2E_ITERS", "12"))
        for i in range(iters):                                  # the split (and the allocator) settle in the first calls
            if i == 6:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            out = layer.forward_many_host(host["prop_feat"], host["prop_mask"], host["tmpl_feat"], host["tmpl_mask"], host["prop_score"],
                                          threads=th, raw_fraction=f)
            res.copy_(out["R"], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (iters - 6)
        print(f"threads={th} raw_fraction={f}: {B/dt:.0f} matches/s  raw={out['raw_problems']} pack_ms={1e3*out['host_pack_seconds']:.2f} est={out['route_estimate']} host_ms={ {k: round(1e3 * v, 2) for k, v in out['host_seconds'].items()} }", flush=True)
