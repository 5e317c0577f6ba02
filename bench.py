#!/usr/bin/env python
"""bench.py -- mask-matches/sec of the DMM-Net matching hot path (cost-build + relaxed-matching solve).

Workload (BASELINE.json configs[1]): N=50 proposals, K=10 templates, 256x448 fp32 soft masks, D=512 features,
20 outer x 5 inner solver iterations, synthetic seeded inputs (dmm_net_b200/synth.py).  One "step" = one pass of the
hot path (cosine K2 -> mask-IoU K1 (+finalize/mix) -> solver+head K3) over a batch of `--batch` independent problems
that is ALREADY RESIDENT in HBM.  The batch (B x 27.5 MB) is far larger than the 126 MB L2, so every step streams
from HBM (no L2 flush needed; stated in config.l2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

N>1 is launched by torchrun (one rank per GPU); ranks shard independent problems (weak scaling, no data-path
collective); time = max over ranks of the CUDA-event time of exactly K steps.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P, O, H, W, D = 50, 10, 256, 448, 512
MAX_ITER, PROJ_ITER, LR, SCORE_W = 20, 5, 0.1, 0.3
MASK_BYTES_PER_MATCH = (P + O) * H * W * 4                       # K1 algorithmic bytes per match: 27,525,120
ALGO_BYTES_PER_MATCH = MASK_BYTES_PER_MATCH + (P + O) * D * 4 + O * P * 4  # 27,650,000
METRIC = "mask-matches/sec (cost-build+Sinkhorn, N=50 K=10 256x448)"
WORKLOAD = "configs[1]: N=50 K=10 256x448 fp32 masks, D=512, 20x5 relaxed-matching iters, cost-build+solve"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self, which):
        """host timestamps of the timed region (taken right after the synchronize on each side)"""
        setattr(self, "t_" + which, time.time())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t_start", None), getattr(self, "t_end", None)
        rows = [r for t, r in self.rows if t0 is not None and t1 is not None and t0 - 0.03 <= t <= t1 + 0.03]
        window = "timed region"
        if not rows:                      # nvidia-smi can take seconds to come up on 8-GPU boxes: fall back, say so
            rows, window = [r for _, r in self.rows], "warm-up + timed region (no sample landed inside the timed region)"
        self.window = window
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def tune_cpu_threads(one):
    """Give the CPU baseline the thread count it runs fastest with: boxes of this pool expose 128 logical CPUs under a
    16-CPU cgroup quota, where torch with 128 threads is several times SLOWER than with 16-32."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        one()
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


CONFIG = {"workload": WORKLOAD, "inputs": "synthetic, seeded (dmm_net_b200/synth.py): fp32 soft masks, N(0,1) features"}


def reference_callable():
    """The CPU arm: one cost-build + solve of the headline problem per call.

    kind "reference": the UNMODIFIED reference code (oracle/_ref = the four hot-path files copied from the reference tree by
    oracle/build_ref.py at build time) -- MatchModel.compute_cost_matrix (match_model.py:49-91) followed by relax_matching
    (relax_match.py:36-105), exactly what MatchModel.forward runs before the assignment-apply.
    kind "port": oracle/match_oracle.py, used only when oracle/_ref is not there."""
    from dmm_net_b200.synth import default_cfg, make_problem
    pr = make_problem(P, O, H, W, D, config=2, index=0)
    try:
        from oracle import build_ref
        have_ref = build_ref.verify()
    except Exception:
        have_ref = False
    if have_ref:
        MatchModel, relax_matching, _ = build_ref.import_reference()
        layer = MatchModel(default_cfg(MAX_ITER, PROJ_ITER, LR, SCORE_W), is_test=1)
        feats = {"proposed": pr.prop_feat, "template": [pr.tmpl_feat]}
        masks = {"proposed": pr.prop_mask, "template": pr.tmpl_mask}
        scores = {"proposal_score": pr.prop_score}

        def one():
            with torch.no_grad():
                sim, _, _, _ = layer.compute_cost_matrix(feats, masks, scores, None)
                return relax_matching(-sim, max_iter=MAX_ITER, proj_iter=PROJ_ITER, lr=LR)[2]
        return one, "reference", "unmodified reference (oracle/_ref: MatchModel.compute_cost_matrix + relax_matching)"
    from oracle import match_oracle as orc

    def one():
        with torch.no_grad():
            sim, _ = orc.cost_matrix(pr.prop_feat, pr.prop_mask, [pr.tmpl_feat], pr.tmpl_mask, SCORE_W, None, expand=True)
            return orc.relax_solve(-sim, MAX_ITER, PROJ_ITER, LR)[2]
    return one, "port", "oracle/match_oracle.py (torch-CPU port of the reference op sequence incl. [O*P,HW] expansion)"


def cpu_reference_rate(seconds_budget: float):
    """The reference's CPU path on this host's cores, timed one problem at a time on a bounded sample."""
    one, kind, what = reference_callable()
    threads = tune_cpu_threads(one)                                   # also warms up
    n, t0 = 0, time.perf_counter()
    while True:
        one()
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds_budget or n >= 200:
            break
    return n / el, n, el, threads, kind, what


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all the host threads it can use, rank 0 only."""
    if rank != 0:
        return
    per_step_budget = max(0.5, min(4.0, 60.0 / max(1, args.steps + args.warmup)))
    one, kind, what = reference_callable()
    threads = tune_cpu_threads(one)
    t0 = time.perf_counter(); one(); t_one = time.perf_counter() - t0
    per_step = max(1, int(per_step_budget / max(t_one, 1e-3)))
    for _ in range(args.warmup):
        for _ in range(per_step):
            one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            one()
    el = time.perf_counter() - t0
    val = args.steps * per_step / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "matches/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG), "run": {"matches_per_step": per_step},
            "cpu_baseline": {"value": val, "unit": "matches/s", "cores": threads, "kind": kind,
                             "sample": f"{args.steps}x{per_step} problems of the headline shape, {what}, {threads} threads"},
            "e2e": {"value": val, "unit": "matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _load_example(name):
    import importlib.util
    ex = os.path.join(ROOT, "examples")
    if ex not in sys.path:
        sys.path.insert(0, ex)
    spec = importlib.util.spec_from_file_location(name, os.path.join(ex, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _shares(op_ms):
    tot = sum(op_ms.values()) or 1.0
    return {k: {"ms": round(v, 3), "share": round(v / tot, 4)} for k, v in sorted(op_ms.items(), key=lambda kv: -kv[1])}


def leg_clip_r50(dev, local_rank):
    """BASELINE configs[2]: ResNet-50 backbone (stock torch) + K5 pooled features -> cosine cost + mask IoU -> solver ->
    apply, on 8-frame clips, N=50 proposals, K=10 templates, 256x448, one GPU.  8 clips advance together (one batch per
    frame); the frames of a clip are sequential."""
    ce = _load_example("synthetic_clip_eval")
    clips, frames = 8, 8
    with ClockSampler(local_rank) as clk:
        ce.clip_eval(clips, 3, 50, 10, (H, W), lazy=False, arch="resnet50", fixed_objects=True)          # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        clk.mark("start")
        # each variant twice, the second one reported: the first pass of a variant still grows the caching allocator (235 MB of
        # pasted masks per frame, a fresh encoder + graph), and a cudaMalloc inside the window is not what the leg measures
        ce.clip_eval(clips, frames + 1, 50, 10, (H, W), lazy=False, arch="resnet50", fixed_objects=True)
        r = ce.clip_eval(clips, frames + 1, 50, 10, (H, W), lazy=False, arch="resnet50", fixed_objects=True, time_ops=True)
        ce.clip_eval(clips, frames + 1, 50, 10, (H, W), lazy=True, arch="resnet50", fixed_objects=True)
        rl = ce.clip_eval(clips, frames + 1, 50, 10, (H, W), lazy=True, arch="resnet50", fixed_objects=True)
        clk.mark("end")
    return {"backbone": "torchvision ResNet-50 + 1x1 necks, eval mode, replayed as one CUDA graph per frame (stock torch.cuda.CUDAGraph)",
            "workload": "configs[2]: ResNet-50 + neck (stock torch) -> K8 paste -> K9 NMS -> K5 pooling -> K2 cosine + K1 IoU -> K3 -> K4 "
                        "-> K6 pyramid + K7 labels; 8 clips x 8 frames, N=50 K=10 256x448, eval.yaml 40x5, 1 GPU",
            "frames_per_s": r["frames"] / (r["ms"] * 1e-3), "ms_per_frame_step": r["ms_per_frame_step"], "clips_in_flight": clips,
            "lazy_pipeline_frames_per_s": rl["frames"] / (rl["ms"] * 1e-3),
            "per_op": _shares(r["op_ms"]), "clocks": clk.summary()}


def leg_eval_r101(dev, local_rank, rank, world):
    """BASELINE configs[3]: ResNet-101, YouTube-VOS-shaped synthetic clips (27 frames, 1..5 objects of F=5, N=50, 255x448),
    clips sharded clip % world == rank, DMM_Model.inference_lazy; no collective in the data path."""
    from dmm_net_b200.sharding import max_over_ranks, shard_indices, sum_over_ranks
    ce = _load_example("synthetic_clip_eval")
    clips_total, frames = 8 * world, 27
    mine = shard_indices(clips_total, rank, world)
    with ClockSampler(local_rank) as clk:
        ce.clip_eval(len(mine), 3, 50, 5, (255, 448), lazy=True, arch="resnet101", seed=4100 + rank)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        clk.mark("start")
        ce.clip_eval(len(mine), 9, 50, 5, (255, 448), lazy=True, arch="resnet101", seed=4200 + rank)      # second warm-up pass: allocator settled
        if world > 1:
            dist.barrier()
        r = ce.clip_eval(len(mine), frames + 1, 50, 5, (255, 448), lazy=True, arch="resnet101", time_ops=True, seed=4000 + rank)
        clk.mark("end")
    ms = max_over_ranks(r["ms"], dev)
    total_frames = sum_over_ranks(r["frames"], dev)
    return {"backbone": "torchvision ResNet-101 + 1x1 necks, eval mode, replayed as one CUDA graph per frame (stock torch.cuda.CUDAGraph)",
            "workload": "configs[3]: ResNet-101 + neck (stock torch), synthetic YouTube-VOS-shaped clips: 27 frames, 1..5 of F=5 objects, "
                        "N=50, 255x448, eval.yaml 40x5; DMM_Model.inference_lazy; clips sharded clip % world == rank, no collective",
            "clips": clips_total, "clips_per_gpu": len(mine), "clips_per_s": clips_total / (ms * 1e-3),
            "frames_per_s": total_frames / (ms * 1e-3), "ms_per_frame_step_max_over_ranks": ms / frames,
            "per_op_rank0": _shares(r["op_ms"]), "scaling": "weak (8 clips per GPU)", "clocks": clk.summary()}


def leg_train(dev, local_rank, rank, world, steps=8):
    """BASELINE configs[4]: train.py-shaped step, random-init ResNet-50, 4 clips x 3 frames per GPU (scripts/train/train_r50.sh),
    matching layer in training mode with autograd, Adam; the data-parallel exchange is ONE all-reduce of all gradients
    (a flat bucket; reference train.py:178-184 DDP, without its redundant per-parameter pass train.py:62-68)."""
    from dmm_net_b200.sharding import max_over_ranks
    ts = _load_example("synthetic_train_step")
    with ClockSampler(local_rank) as clk:
        clk.mark("start")
        r = ts.train_loop("resnet50", 4, 3, 3, 50, (H, W), steps=steps, warmup=3, reduce="flat" if world > 1 else "none")
        clk.mark("end")
        ddp = ts.train_loop("resnet50", 4, 3, 3, 50, (H, W), steps=4, warmup=2, reduce="ddp") if world > 1 else None
    step_ms = max_over_ranks(r["step_ms"], dev)
    red_ms = max_over_ranks(r["reduce_ms"], dev)
    skew_ms = max_over_ranks(r["skew_ms"], dev)
    out = {"workload": "configs[4]: train.py-shaped step, ResNet-50 + neck + conv decoder (stock torch, random init), 4 clips x 3 frames "
                       "per GPU, N=50 F=3 256x448, train.yaml 10x5, K8/K5/K2/K1/K3/K4/K6 with autograd, fused Adam",
           "step_ms": step_ms, "host_ms": r["host_ms"], "fwd_ms": r["fwd_ms"], "bwd_ms": r["bwd_ms"], "opt_ms": r["opt_ms"],
           "clips_per_s": world * 4 / (step_ms * 1e-3), "grad_bytes": r["grad_bytes"], "grad_tensors": r["n_grad_tensors"],
           "exchange": r["reduce"], "exposed_allreduce_ms": red_ms if world > 1 else 0.0,
           "rank_skew_wait_ms": skew_ms if world > 1 else 0.0,
           "allreduce": None, "scaling": "weak (4 clips per GPU)", "clocks": clk.summary(),
           "note": "the step is bound by the host thread launching the stock-torch backbone kernels (host_ms ~ step_ms); the all-reduce "
                   "runs after backward, fully exposed, and is timed by CUDA events around it; a 4-byte all-reduce right before it absorbs "
                   "the wait for the slowest rank (rank_skew_wait_ms), so exposed_allreduce_ms is the transfer itself"}
    if world > 1:
        bus = 2 * (world - 1) / world * r["grad_bytes"] / (red_ms * 1e-3) / 1e9 if red_ms > 0 else None
        out["allreduce"] = {"collective": "NCCL all-reduce (AVG), one flat fp32 bucket", "bytes": r["grad_bytes"], "ms": red_ms,
                            "bus_gbs": bus, "share_of_step": red_ms / step_ms}
        out["torch_ddp_step_ms"] = max_over_ranks(ddp["step_ms"], dev)
    return out


def leg_worker(name):
    """`bench.py --leg-worker NAME`: one secondary leg in its OWN process group (spawned by rank 0 of the main run, under
    torchrun when N > 1), so that nothing a leg does -- another model, another allocator state, a CUDA fault -- can cost the
    headline line.  Rank 0 prints one JSON line tagged "leg"."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    if name == "clip":
        out = leg_clip_r50(dev, local_rank)
    elif name == "eval":
        out = leg_eval_r101(dev, local_rank, rank, world)
    elif name == "train":
        out = leg_train(dev, local_rank, rank, world)
    else:
        raise SystemExit(f"unknown leg {name}")
    out["wall_s"] = round(time.perf_counter() - t0, 1)
    if rank == 0:
        print(json.dumps({"leg": name, "result": out}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_leg(name, world, index):
    """Rank 0 of the main run: spawn the leg's process group on the same GPUs and read its JSON line back."""
    cmd = [sys.executable]
    if world > 1:
        port = int(os.environ.get("MASTER_PORT", "29500")) + 101 + index
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                "--master-port", str(port)]
    cmd += [os.path.join(ROOT, "bench.py"), "--leg-worker", name, "--gpus", str(world)]
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "ROLE_NAME", "ROLE_WORLD_SIZE",
                        "GROUP_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT") and not k.startswith("TORCHELASTIC_")}
    import signal
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=600)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)                      # the leg's own session: torchrun and every rank under it
        except OSError:
            pass
        proc.communicate()
        return {"error": "leg timed out after 600 s (its process group was killed)"}
    for line in reversed(out.splitlines()):
        if line.startswith('{"leg"'):
            return json.loads(line)["result"]
    tail = [x for x in err.splitlines() if x.strip() and "Warning" not in x][-6:] + \
           [x for x in out.splitlines() if x.startswith("libdmm_b200")][:3]
    return {"error": f"leg exited with code {proc.returncode}", "stderr_tail": tail}


def host_rooflines(host_masks, dev, threads):
    """What bounds the host-buffer entry: every mask byte has to leave host memory once, through the packing cores (a
    streaming read) or through the copy engine (pinned H2D).  Measured here, on the buffers the e2e leg uses: each route
    alone, and BOTH AT ONCE (the two routes share the host's memory system, so their concurrent rates -- not the sum of
    the solo rates -- are the ceiling of the two-route scheme).  Returns GB/s: (read_solo, h2d_solo, read_conc, h2d_conc)."""
    import ctypes
    from dmm_net_b200 import _lib
    lib = _lib.load()
    nbytes = host_masks.numel() * 4

    def read_rate(reps):
        gbs = ctypes.c_double(0.0)
        lib.dmm_host_read_bandwidth(ctypes.c_void_p(host_masks.data_ptr()), nbytes, int(threads), reps, ctypes.byref(gbs))
        return gbs.value

    dst = torch.empty_like(host_masks, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(host_masks, non_blocking=True)
    torch.cuda.synchronize()
    read_solo = read_rate(3)
    a.record()
    for _ in range(2):
        dst.copy_(host_masks, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    h2d_solo = 2 * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9
    # both at once: queue enough copies to outlast the read passes, read while they run
    ncopy = max(2, int(3.5 * h2d_solo / max(read_solo, 1.0)) + 2)
    a.record()
    for _ in range(ncopy):
        dst.copy_(host_masks, non_blocking=True)
    b.record()
    read_conc = read_rate(3)
    torch.cuda.synchronize()
    h2d_conc = ncopy * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9
    return read_solo, h2d_solo, read_conc, h2d_conc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024, help="independent problems per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg (rank 0, N=1)")
    ap.add_argument("--e2e-steps", type=int, default=40, help="timed host-buffer steps (the pipeline drain at the end is amortised over them)")
    ap.add_argument("--secondary-steps", type=int, default=5, help="timed steps of the full-layer secondary metric (0: skip)")
    ap.add_argument("--e2e-threads", type=int, default=0, help="host threads for mask packing (0: cgroup-aware default)")
    ap.add_argument("--legs", default="clip,eval,train", help="secondary legs (BASELINE configs[2..4]): comma list of clip,eval,train; '' = none")
    ap.add_argument("--leg-worker", default="", help="internal: run one secondary leg in this process group and print its JSON")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.leg_worker:
        leg_worker(args.leg_worker)
        return

    import torch.distributed as dist
    from dmm_net_b200 import ops
    from dmm_net_b200.synth import make_problems
    assert torch.cuda.is_available(), "bench.py needs a GPU (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    clk = ClockSampler(local_rank)
    clk.__enter__()                                                              # up before the data exists: slow to start
    pr = make_problems(B, P, O, H, W, D, seed=2000 + rank, device=dev)          # synthetic, generated on the device
    launches = 0

    chunks_seen = []

    def hot_path(k1_events=None):
        """cost-build + solve for the resident batch through the production inference path (ops.cost_and_solve: K2 ->
        K1(+finalize) -> K3 per chunk, chunks staggered on two streams); returns the mean-iterate assignment R."""
        nonlocal launches
        ev = [] if k1_events is None else k1_events
        out = ops.cost_and_solve(pr.prop_feat, pr.prop_mask, pr.tmpl_feat, pr.tmpl_mask, pr.prop_score, max_iter=MAX_ITER,
                                 proj_iter=PROJ_ITER, lr=LR, score_weight=SCORE_W, is_test=True, k1_events=ev)
        launches += 4 * len(ev) if k1_events is not None else 0
        chunks_seen.append(len(ev))
        return out["R"]

    with torch.no_grad():
        for _ in range(args.warmup):
            hot_path()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        launches = 0
        ev = [[] for _ in range(args.steps)]
        t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        clk.mark("start")
        t_beg.record()
        for k in range(args.steps):
            R = hot_path(ev[k])
        t_end.record()
        torch.cuda.synchronize()
        clk.mark("end")
        clk.__exit__()
        ms_total = t_beg.elapsed_time(t_end)
        k1_ms = sum(a.elapsed_time(b) for step_ev in ev for a, b in step_ev) / args.steps   # sum over the step's K1 launches
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = t.item()
        dist.barrier()
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e: same metric through the reference-facing module with HOST (pinned) buffers -------------------
    from dmm_net_b200.modules.match_model import MatchModel
    from dmm_net_b200.synth import default_cfg
    layer = MatchModel(default_cfg(MAX_ITER, PROJ_ITER, LR, SCORE_W), is_test=1)
    Be = min(B, 64)
    from dmm_net_b200 import hostmem
    host, host_node = {}, None
    for k in ("prop_feat", "prop_mask", "tmpl_feat", "tmpl_mask", "prop_score"):     # pinned, on the GPU's NUMA node when allowed
        host[k], host_node = hostmem.pinned_near_gpu(getattr(pr, k)[:Be].cpu(), local_rank)
    res_host = torch.empty(Be, O, max(P, O + 1), pin_memory=True)
    res_ms = torch.empty(Be, O, pin_memory=True)
    e2e_info = {}

    def e2e_step():
        # the reference-facing call with HOST buffers: host cores bit-pack the masks, bits+features cross PCIe,
        # K2 -> K1(packed) -> K3 on the device, the assignment and scores come back to pinned host memory
        out = layer.forward_many_host(host["prop_feat"], host["prop_mask"], host["tmpl_feat"], host["tmpl_mask"],
                                      host["prop_score"], device=dev,
                                      threads=args.e2e_threads or max(2, ops.host_threads() // world))
        res_host.copy_(out["R"], non_blocking=True)
        res_ms.copy_(out["match_score"], non_blocking=True)
        e2e_info.update(h2d=out["h2d_bytes"], packed=out["host_packed_bytes"], threads=out["host_threads"],
                        raw=out["raw_problems"], pack_s=out["host_pack_seconds"], est=out["route_estimate"])

    with torch.no_grad():
        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(6):                                            # pinned staging buffers exist now; the raw/packed split
            e2e_step()                                                #   settles on the measured route speeds
            torch.cuda.synchronize()
        t_wall = time.perf_counter()
        a.record()
        for _ in range(args.e2e_steps):
            e2e_step()
        b.record()
        torch.cuda.synchronize()
        # host work (mask packing) sits inside the region: take the larger of the device-event span and the wall clock
        e2e_ms = max(a.elapsed_time(b), (time.perf_counter() - t_wall) * 1e3)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_val = world * Be * args.e2e_steps / (e2e_ms * 1e-3)
    d2h = (res_host.numel() + res_ms.numel()) * 4

    # ---- what bounds e2e: host-DRAM streaming read (packer's thread team) and the pinned H2D copy rate, measured here -----
    e2e_threads = args.e2e_threads or max(2, ops.host_threads() // world)
    if world > 1:
        dist.barrier()                                                # all ranks measure at the same time: aggregate rates
    try:
        host_read_gbs, h2d_gbs, read_conc, h2d_conc = host_rooflines(host["prop_mask"], dev, e2e_threads)
    except Exception as exc:                                          # a probe must never cost the headline line
        print(f"host_rooflines failed: {type(exc).__name__}: {exc}", file=sys.stderr)
        host_read_gbs = h2d_gbs = read_conc = h2d_conc = 0.0
    if world > 1:
        t = torch.tensor([host_read_gbs, h2d_gbs, read_conc, h2d_conc], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        host_read_gbs, h2d_gbs, read_conc, h2d_conc = t.tolist()
    e2e_gbs = e2e_val * MASK_BYTES_PER_MATCH / 1e9                    # mask bytes that left host DRAM per second, whole job

    # ---- secondary: the full layer incl. assignment-apply (K4), resident inputs, this rank only ------
    secondary = None
    if rank == 0 and args.secondary_steps > 0:
        with torch.no_grad():
            full = lambda: layer.forward_many(pr.prop_feat, pr.prop_mask, pr.tmpl_feat, pr.tmpl_mask, pr.prop_score)["full_outmask"]
            for _ in range(2):
                out_full = full()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(args.secondary_steps):
                out_full = full()
            b.record()
            torch.cuda.synchronize()
            ms_full = a.elapsed_time(b) / args.secondary_steps
        secondary = {"full_layer": {"matches_per_s_per_gpu": B / (ms_full * 1e-3), "ms_per_step": ms_full,
                                    "what": "MatchModel.forward_many incl. K4 assignment-apply writing full_outmask [B,O,H,W], resident inputs"}}
        del out_full
    h2d = e2e_info.get("h2d", 0)

    # ---- secondary: the layer in TRAINING mode (train.yaml 10x5, targets given -> match loss), forward + backward, resident -----
    if rank == 0 and args.secondary_steps > 0:
        try:
            layer_tr = MatchModel(default_cfg(10, 5, LR, SCORE_W), is_test=0)
            targets = (pr.tmpl_mask > 0.3).float()
            pf = pr.prop_feat.clone().requires_grad_(True)
            tf = pr.tmpl_feat.clone().requires_grad_(True)

            def train_step():
                out = layer_tr.forward_many(pf, pr.prop_mask, tf, pr.tmpl_mask, pr.prop_score, targets)
                (out["full_outmask"].mean() + out["match_score"].sum() + out["cost_loss"].sum()).backward()
                pf.grad = tf.grad = None

            train_step()
            timer = ops.KernelTimer()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ops.set_kernel_timer(timer)
            a.record()
            for _ in range(3):
                train_step()
            b.record()
            ops.set_kernel_timer(None)
            op_ms = {k: v / 3 for k, v in timer.totals().items()}
            ms_tr = a.elapsed_time(b) / 3
            k1_tr = op_ms.get("K1 mask_iou", 0.0)
            bytes_tr = (P + 2 * O) * H * W * 4 * B                       # proposals + previous masks + targets, each row once
            peak_hbm, _ = measured_peaks()
            secondary["train_layer"] = {
                "what": "MatchModel.forward_many in training mode (targets -> match loss; K2, K1 with the targets as second template set in "
                        "ONE pass over 70 rows, K3 fwd+bwd, K4 fwd+bwd, K2 bwd), forward + backward, resident inputs",
                "matches_per_s_per_gpu": B / (ms_tr * 1e-3), "ms_per_step_fwd_bwd": ms_tr, "forward_op_ms": {k: round(v, 3) for k, v in op_ms.items()},
                "k1_roofline": {"algorithmic_bytes_per_launch": bytes_tr, "kernel_ms_incl_finalize": k1_tr,
                                "achieved_gbs": bytes_tr / (k1_tr * 1e-3) / 1e9 if k1_tr else None,
                                "frac": bytes_tr / (k1_tr * 1e-3) / 1e9 / peak_hbm if k1_tr else None}}
            del targets, pf, tf
        except Exception as exc:
            secondary["train_layer"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- secondary legs: BASELINE configs[2], [3], [4] -- each in its own process group on the same GPUs (see leg_worker);
    #      the ranks of this run free their memory and wait on the host (gloo), not on the GPU
    del pr, R
    torch.cuda.empty_cache()
    legs = [x for x in args.legs.split(",") if x and not (x == "clip" and world > 1)]
    leg_out = {}
    import datetime
    del host
    ctl = None
    if world > 1 and legs:
        try:                                                          # host-side waiting: the idle ranks must not spin on the GPU
            ctl = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=45))
        except Exception as exc:
            print(f"gloo control group unavailable ({type(exc).__name__}: {exc}); the ranks wait on an NCCL barrier instead", file=sys.stderr)
    if legs:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=ctl) if ctl is not None else dist.barrier()
        if rank == 0:
            names = {"clip": "clip_r50", "eval": "eval_r101", "train": "train"}
            for i, name in enumerate(legs):
                leg_out[names.get(name, name)] = run_leg(name, world, i)
        if world > 1:
            dist.barrier(group=ctl) if ctl is not None else dist.barrier()
    if secondary is None:
        secondary = {}
    secondary.update(leg_out)

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = MASK_BYTES_PER_MATCH * B / (k1_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "matches/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG),
            "run": {"problems_per_step_per_gpu": B, "k1_launches_per_step": chunks_seen[-1] if chunks_seen else None,
                    "parallelism": f"dp{world} (problems sharded, no collective)",
                    "l2": f"inputs {B * MASK_BYTES_PER_MATCH / 1e9:.2f} GB per GPU >> 126 MB L2, no flush needed",
                    "full_step_gbs_per_gpu": ALGO_BYTES_PER_MATCH * B / (ms_total / args.steps * 1e-3) / 1e9},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": "matches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "problems_per_step": Be, "host_bytes_packed_per_step": e2e_info.get("packed"),
                    "host_threads": e2e_info.get("threads"), "problems_sent_as_fp32": e2e_info.get("raw"),
                    "host_buffers_numa_node": host_node,
                    "host_pack_ms_per_step": None if e2e_info.get("pack_s") is None else 1e3 * e2e_info["pack_s"],
                    "route_seconds_per_problem(pack,dma)": e2e_info.get("est"),
                    "roofline": {"bound": "host: every mask byte leaves host memory once, through the packing cores or through the copy engine",
                                 "achieved_gbs": e2e_gbs, "host_dram_read_peak_gbs": host_read_gbs,
                                 "pinned_h2d_peak_gbs": h2d_gbs,
                                 "concurrent_read_gbs": read_conc, "concurrent_h2d_gbs": h2d_conc,
                                 "ceiling_gbs": read_conc + h2d_conc,
                                 "frac_of_ceiling": e2e_gbs / (read_conc + h2d_conc) if read_conc + h2d_conc else None,
                                 "ceiling_matches_per_s": (read_conc + h2d_conc) * 1e9 / MASK_BYTES_PER_MATCH,
                                 "ceiling_model": "packed route <= what the packing threads stream from host memory, raw route <= the pinned H2D "
                                                  "rate of the copy engine, BOTH MEASURED WHILE THE OTHER RUNS (they share the host's memory "
                                                  "system); the ceiling is the sum of the concurrent rates, over all ranks",
                                 "how": f"dmm_host_read_bandwidth: {e2e_threads} threads per rank streaming the pinned proposal-mask buffer "
                                        "(best of 3); H2D: pinned copies of the same buffer, CUDA events; solo and concurrently; summed over ranks"},
                    "api": "MatchModel.forward_many_host: pinned host fp32 inputs; the host cores bit-pack most masks (bits "
                           "cross PCIe) while the copy engine DMAs the rest as fp32; features+scores H2D, R and match_score D2H"},
            "gpu_launches": launches,
            "parity": {"headline_path": "pinned: golden vectors generated from the unmodified reference (tests/golden, oracle/make_golden*.py); "
                                        "IoU / greedy init / selection bit-exact, floats <= 1e-4, arg-max equal; cpu arm = the reference itself",
                       "unpinned_by_necessity": ["K5 / K5-TC proposal-feature pooling: ROIAlign lives in the un-vendored maskrcnn_benchmark fork; "
                                                 "stand-in oracle torchvision roi_align(aligned=False) (secondary legs clip_r50 / eval_r101 / train use it)",
                                                 "K9 box NMS: same fork; restated from the published algorithm, property-tested"]},
            "secondary": secondary,
            "roofline": {"bound": "hbm", "kernel": "mask_iou_partial_kernel(+finalize)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": MASK_BYTES_PER_MATCH * B, "kernel_ms": k1_ms},
        }
        if world == 1 and args.cpu_seconds > 0:
            rate, n, el, threads, kind, what = cpu_reference_rate(args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": "matches/s", "cores": threads, "kind": kind,
                                    "sample": f"{n} problems of the headline shape in {el:.1f} s, {what}"}
        prof = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(prof):                                      # an ncu number (dram bytes per match of one capture): only
            try:                                                      #   valid for the kernel source it was captured from
                import hashlib
                rec = json.load(open(prof))
                src_sha = hashlib.sha256(open(os.path.join(ROOT, "dmm_net_b200", "csrc", "mask_iou.cu"), "rb").read()).hexdigest()
                if rec.get("source_sha256") in (None, src_sha):
                    line["roofline"]["traffic"] = rec.get("traffic_bytes_per_match", 0) * B or None
                    line["roofline"]["traffic_source"] = rec.get("source")
                else:
                    line["roofline"]["traffic_source"] = "stale: mask_iou.cu changed since the ncu capture in profiles/k1_traffic.json"
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
