#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
for v in T512 T256 S1024 S512; do
  GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
done
for v in base T512 S512; do
  if [ $v = base ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --steps 80 2>/dev/null | tail -1 > gpurun_out/c16_bench_$v.json
  python -c "
import json; d=json.load(open('gpurun_out/c16_bench_$v.json')); k=d['kernels']
print('$v', 'views/s', d['value'], 'batch ms', sorted(d['batches']['device_ms'])[:3], 'seq', d['sequential_ms_per_view'], 'fwd', k['gsb_composite_fwd']['avg_ms'], 'bwd', k['gsb_composite_bwd']['avg_ms'], 'shade_bwd', k['gsb_shade_bwd']['avg_ms'])"
done
