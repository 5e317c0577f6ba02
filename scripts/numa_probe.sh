#!/bin/bash
# host topology of the GPU box: NUMA nodes, allowed cpus / mems, each GPU's node
echo "nodes online: $(cat /sys/devices/system/node/online)"
for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4,$5}')"; done
echo "cpuset.cpus.effective: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"
echo "cpuset.mems.effective: $(cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null)"
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"
echo "affinity: $(python -c 'import os; print(sorted(os.sched_getaffinity(0)))' | cut -c1-200)"
nvidia-smi topo -m | head -12
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from dmm_net_b200 import hostmem
for i in range(torch.cuda.device_count()):
    print("gpu", i, "numa", hostmem.gpu_numa_node(i))
print("mem nodes", hostmem.memory_nodes())
for n in hostmem.memory_nodes():
    with hostmem.prefer_node(n) as ok:
        print("prefer", n, ok)
PY
lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" | head -10
