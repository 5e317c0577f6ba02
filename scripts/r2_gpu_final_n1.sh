#!/bin/bash
# final single-GPU record of the round: tests, bench (driver's command line), reference arm, launch list of the bench itself
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
grep -v Warning gpurun_out/r2_bench_n1.err | tail -3
python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mask_iou|cosine|relax_solve|assign_apply|mask_pack" -c 60 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --e2e-steps 1 --secondary-steps 1 --legs '' > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_bench_launches.csv | cut -d, -f5,15- | sort | uniq -c | sort -rn | head -5
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2_bench_n1.json") if x.startswith("{")]
d=json.loads(l[-1])
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"], d["roofline"].get("traffic"), d["clocks"])
e=d["e2e"]; print("e2e", e["value"], json.dumps(e["roofline"])[:900])
print(d["cpu_baseline"])
for leg in ("clip_r50","eval_r101","train","train_layer","full_layer"):
    v=d["secondary"].get(leg,{})
    print(leg, {k:v.get(k) for k in ("frames_per_s","lazy_pipeline_frames_per_s","clips_per_s","step_ms","matches_per_s_per_gpu","wall_s","error")})
r=[x for x in open("gpurun_out/r2_bench_reference_arm.json") if x.startswith("{")]
print(json.loads(r[-1])["value"], json.loads(r[-1])["cpu_baseline"]["kind"])
PY
