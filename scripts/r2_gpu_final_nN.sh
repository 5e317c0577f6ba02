#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"
grep -v Warning gpurun_out/r2_bench_n$N.err | tail -3
grep -h "libdmm" gpurun_out/r2_bench_n$N.json gpurun_out/r2_bench_n$N.err | head -3
python - <<PY
import json
l=[x for x in open("gpurun_out/r2_bench_n$N.json") if x.startswith("{")]
d=json.loads(l[-1])
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"], d["clocks"])
e=d["e2e"]; r=e["roofline"]; print("e2e", e["value"], {k:r[k] for k in ("achieved_gbs","host_dram_read_peak_gbs","pinned_h2d_peak_gbs","concurrent_read_gbs","concurrent_h2d_gbs","frac_of_ceiling")})
for leg in ("eval_r101","train"):
    v=d["secondary"].get(leg,{})
    print(leg, {k:v.get(k) for k in ("frames_per_s","clips_per_s","step_ms","host_ms","exposed_allreduce_ms","rank_skew_wait_ms","torch_ddp_step_ms","wall_s","error","stderr_tail")}, v.get("allreduce"))
PY
