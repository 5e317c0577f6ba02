#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
for leg in ${2:-clip eval}; do
if [ "$N" = "1" ]; then python bench.py --leg-worker $leg --gpus 1 > gpurun_out/r2_leg_${leg}_n$N.json 2> gpurun_out/r2_leg_${leg}_n$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --leg-worker $leg --gpus $N > gpurun_out/r2_leg_${leg}_n$N.json 2> gpurun_out/r2_leg_${leg}_n$N.err; fi
echo "$leg rc=$?"; grep -v Warning gpurun_out/r2_leg_${leg}_n$N.err | tail -4
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2_leg_${leg}_n$N.json") if x.startswith('{"leg"')]
    d=json.loads(l[-1])["result"]
    print({k:d.get(k) for k in ("frames_per_s","lazy_pipeline_frames_per_s","clips_per_s","ms_per_frame_step","ms_per_frame_step_max_over_ranks","wall_s")})
    for k,x in list(d.get("per_op", d.get("per_op_rank0", {})).items())[:7]: print("   ",k,x)
except Exception as e: print("no result", e)
PY
done
