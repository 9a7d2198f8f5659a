#!/bin/bash
for m in full noup nodown; do
  GSB_E2E_DEBUG=$m python bench.py --no-cpu-baseline --no-configs --no-train-step --steps 48 2>/dev/null | tail -1 > gpurun_out/c41_$m.json
  python -c "
import json; d=json.load(open('gpurun_out/c41_$m.json')); print('$m', d['value'], d['e2e']['value'], d['e2e']['steps'])"
done
