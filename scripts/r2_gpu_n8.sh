#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
bash scripts/numa_probe.sh > gpurun_out/r2_numa_probe_n$N.txt 2>&1
cat gpurun_out/r2_numa_probe_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 20 --legs ${2:-train} > gpurun_out/r2_bench_n${N}_short.json 2> gpurun_out/r2_bench_n${N}_short.err
echo rc=$?
grep -v Warning gpurun_out/r2_bench_n${N}_short.err | tail -5
python - <<PY
import json
l=[x for x in open("gpurun_out/r2_bench_n${N}_short.json") if x.startswith("{")]
d=json.loads(l[-1])
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"])
print("e2e",json.dumps(d["e2e"],indent=1)[:2500])
print(json.dumps({k:v for k,v in d["secondary"].items() if k!="full_layer"},indent=1)[:3000])
PY
