#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/c54_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c54_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-train-step --steps 80 2>/dev/null | tail -1 > gpurun_out/c54_bench.json
python -c "
import json; d=json.load(open('gpurun_out/c54_bench.json')); print('views/s', d['value'], {k: (v.get('views_per_s') or v.get('D14_ms_per_view')) for k, v in d.get('configs', {}).items()})"
GSB_BIN2_RADIX=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-train-step --steps 16 2>/dev/null | tail -1 > gpurun_out/c54_bench_radix.json
python -c "
import json; d=json.load(open('gpurun_out/c54_bench_radix.json')); print('radix', d['value'], {k: (v.get('views_per_s') or v.get('D14_ms_per_view')) for k, v in d.get('configs', {}).items()})"
