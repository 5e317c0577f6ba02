#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests/test_gpu_parity.py -q -k "host_buffer" 2>&1 | tail -2
if [ "$N" = "1" ]; then python bench.py --steps 5 --warmup 3 --cpu-seconds 0 --legs '' --secondary-steps 0 > gpurun_out/r2_e2e_n$N.json 2> gpurun_out/r2_e2e_n$N.err
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 5 --warmup 3 --cpu-seconds 0 --legs '' --secondary-steps 0 > gpurun_out/r2_e2e_n$N.json 2> gpurun_out/r2_e2e_n$N.err; fi
echo rc=$?; grep -v Warning gpurun_out/r2_e2e_n$N.err | tail -3
python - <<PY
import json
l=[x for x in open("gpurun_out/r2_e2e_n$N.json") if x.startswith("{")]
d=json.loads(l[-1]); e=d["e2e"]
print("value", d["value"], "e2e", e["value"], {k:e[k] for k in ("host_threads","problems_sent_as_fp32","host_pack_ms_per_step","route_seconds_per_problem(pack,dma)")}, e["roofline"]["frac_of_ceiling"], e["roofline"]["ceiling_gbs"])
PY
