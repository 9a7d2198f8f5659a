#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_shade_gpu.py tests/test_splat_gpu.py tests/test_parity_fullsize_gpu.py tests/test_prefilter_gpu.py -q -x > gpurun_out/c46_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c46_tests.log
timeout 100 python scripts/bench_composite.py --iters 24 | tail -1
timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --no-train-step --steps 80 2>/dev/null | tail -1 > gpurun_out/c46_bench.json
python -c "
import json; d=json.load(open('gpurun_out/c46_bench.json')); k=d['kernels']
print('views/s', d['value'], 'batch ms', sorted(d['batches']['device_ms'])[:3], {n: v['avg_ms'] for n, v in k.items() if 'shade' in n})"
