#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/c39_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c39_tests.log
timeout 100 python scripts/bench_composite.py --iters 24 | tail -1
timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --steps 80 2>/dev/null | tail -1 > gpurun_out/c39_bench.json
python -c "
import json; d=json.load(open('gpurun_out/c39_bench.json')); k=d['kernels']
print('views/s', d['value'], 'batch ms', sorted(d['batches']['device_ms'])[:3], 'seq', d['sequential_ms_per_view'], {n: v['avg_ms'] for n, v in k.items()})
print('parity', json.dumps(d.get('parity'))[:400])"
