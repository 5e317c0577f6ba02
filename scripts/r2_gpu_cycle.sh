#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_exit_parity.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.txt
tail -12 gpurun_out/r2_gputests.txt
python scripts/prof_k5.py tc 2>&1 | grep "bwd\|64 frames x 50 ROIs (typical)"
python bench.py --steps 10 --warmup 3 --cpu-seconds 4 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
grep -v Warning gpurun_out/r2_bench_n1.err | tail -3
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2_bench_n1.json") if x.startswith("{")]
d=json.loads(l[-1])
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"], d["roofline"].get("traffic"))
e=d["e2e"]; print("e2e", e["value"], e["roofline"]["frac_of_ceiling"], e["host_pack_ms_per_step"], e["route_seconds_per_problem(pack,dma)"])
for leg in ("clip_r50","eval_r101"):
    v=d["secondary"][leg]
    print(leg, v.get("wall_s"), v.get("frames_per_s"), v.get("lazy_pipeline_frames_per_s"), v.get("error"))
    for k,x in list(v.get("per_op", v.get("per_op_rank0", {})).items())[:6]: print("   ",k,x)
print("train", {k:d["secondary"]["train"].get(k) for k in ("step_ms","host_ms","wall_s","error")})
print("train_layer", json.dumps(d["secondary"]["train_layer"])[:600])
PY
