#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --steps 10 --warmup 3 --cpu-seconds 4 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
fi
echo rc=$?
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
l=[x for x in open("gpurun_out/r2_bench_n$N.json") if x.startswith("{")]
d=json.loads(l[-1])
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"])
print("e2e",json.dumps(d["e2e"],indent=1)[:1500])
print(json.dumps(d["secondary"],indent=1)[:6000])
PY
