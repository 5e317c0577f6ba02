"""debug: the train loop with the data seeds the 8 ranks of an 8-GPU run would use, on one GPU"""
import os, sys, importlib.util
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
spec = importlib.util.spec_from_file_location("ts", os.path.join(os.path.dirname(__file__), "..", "examples", "synthetic_train_step.py"))
ts = importlib.util.module_from_spec(spec); spec.loader.exec_module(ts)
torch.cuda.set_device(0)
for off in [int(a) for a in sys.argv[1:]] or range(8):
    r = ts.train_loop("resnet50", 4, 3, 3, 50, (256, 448), steps=11, warmup=0, reduce="none", seed_offset=off)
    torch.cuda.synchronize()
    print("seed offset", off, "ok: step", round(r["step_ms"], 1), "ms loss", r["hist"][-1], flush=True)
