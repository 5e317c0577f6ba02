#!/bin/bash
# N-GPU: the train leg worker three times (hunting a rare fault), then the short bench with all legs through subprocesses
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
for i in 1 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+i)) bench.py --leg-worker train --gpus $N > gpurun_out/r2_trainleg_n${N}_$i.json 2> gpurun_out/r2_trainleg_n${N}_$i.err
  echo "train leg run $i rc=$?"; grep -h "libdmm_b200\|timed out" gpurun_out/r2_trainleg_n${N}_$i.err | sort | uniq -c | head -5
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r2_trainleg_n${N}_$i.json") if x.startswith('{"leg"')]
    d=json.loads(l[-1])["result"]
    print({k:d[k] for k in ("step_ms","host_ms","exposed_allreduce_ms","rank_skew_wait_ms","torch_ddp_step_ms","allreduce") if k in d})
except Exception as e: print("no result", e)
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 20 --legs ${2:-eval,train} > gpurun_out/r2_bench_n${N}_short.json 2> gpurun_out/r2_bench_n${N}_short.err
echo bench rc=$?
grep -v Warning gpurun_out/r2_bench_n${N}_short.err | tail -5
python - <<PY
import json
l=[x for x in open("gpurun_out/r2_bench_n${N}_short.json") if x.startswith("{")]
d=json.loads(l[-1])
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"])
e=d["e2e"]; print("e2e",e["value"], json.dumps(e["roofline"])[:600])
print(json.dumps({k:v for k,v in d["secondary"].items() if k!="full_layer"},indent=1)[:3500])
PY
