#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for ns in 3 4 5 6 8; do
  python bench.py --no-cpu-baseline --no-e2e --no-configs --steps 80 --streams $ns 2>/dev/null | tail -1 > gpurun_out/c21_sweep_$ns.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/c21_sweep_$ns.json')); print('streams', $ns, d['value'], d['ms_per_step'], sorted(d['batches']['device_ms'])[:4])"
done
for v in C2 C3 C4M4; do
  GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so python bench.py --no-cpu-baseline --no-e2e --no-configs --steps 80 2>/dev/null | tail -1 > gpurun_out/c21_$v.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/c21_$v.json')); print('$v', d['value'], d['ms_per_step'], sorted(d['batches']['device_ms'])[:4])"
done
