#!/bin/bash
for c in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c python bench.py --no-cpu-baseline --no-configs --no-train-step --steps 48 2>/dev/null | tail -1 > gpurun_out/c42_$c.json
  python -c "
import json; d=json.load(open('gpurun_out/c42_$c.json')); print('connections $c', d['value'], d['e2e']['value'], d['e2e']['steps'])"
done
