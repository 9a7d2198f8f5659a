#!/bin/bash
# host topology of the GPU box: NUMA nodes, which node each GPU hangs off, the CPUs this container may use
lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"
nproc; taskset -p $$
nvidia-smi topo -m 2>&1 | head -30
for d in /sys/bus/pci/devices/*; do
  if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi
done
python - <<'PY'
import pynvml, os
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    words = (os.cpu_count() + 63) // 64
    try:
        m = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, x in enumerate(m) for b in range(64) if (x >> b) & 1]
        print(i, pynvml.nvmlDeviceGetPciInfo(h).busId, 'ideal cpus', cpus[:4], '...', cpus[-4:], len(cpus))
    except Exception as e:
        print(i, 'affinity failed', e)
print('allowed', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])
PY
cat /sys/devices/system/node/node*/cpulist 2>/dev/null
numactl -H 2>/dev/null | head -12
