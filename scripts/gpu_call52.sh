#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_splat_gpu.py tests/test_parity_fullsize_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/c52_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c52_tests.log
for m in part radix; do
  if [ $m = radix ]; then export GSB_BIN2_RADIX=1; else unset GSB_BIN2_RADIX; fi
  timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --no-train-step --steps 80 2>/dev/null | tail -1 > gpurun_out/c52_$m.json
  python -c "
import json; d=json.load(open('gpurun_out/c52_$m.json'))
print('$m', 'views/s', d['value'], 'batch ms', sorted(d['batches']['device_ms'])[:3], 'parity ids', (d.get('parity') or {}).get('ids_equal'))"
done
unset GSB_BIN2_RADIX
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tile_count|tile_prefix|tile_offsets|tile_scatter" -c 8 python bench.py --no-cpu-baseline --no-configs --no-e2e --no-train-step --steps 8 2>/dev/null | grep -E "tile_|isect_tiles|gpu__time" | head -24
