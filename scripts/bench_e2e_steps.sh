#!/bin/bash
# bench.py's e2e leg with different numbers of timed host-buffer steps (same box, same process layout)
for n in 5 20 60; do
  python bench.py --cpu-seconds 0 --secondary-steps 0 --steps 5 --e2e-steps $n 2>/dev/null > /tmp/b_$n.json
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
d = json.load(open(f"/tmp/b_{n}.json"))
e = d["e2e"]
print("e2e_steps", n, round(e["value"]), "raw", e["problems_sent_as_fp32"], "pack_ms", round(e["host_pack_ms_per_step"], 2))
PY
done
