#!/bin/bash
for v in default TP8; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --no-train-step --steps 80 2>/dev/null | tail -1 > gpurun_out/c53_$v.json
python -c "
import json; d=json.load(open('gpurun_out/c53_$v.json')); print('$v', 'views/s', d['value'], sorted(d['batches']['device_ms'])[:3])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tile_count|tile_prefix|tile_offsets|tile_scatter" -c 4 python bench.py --no-cpu-baseline --no-configs --no-e2e --no-train-step --steps 8 2>/dev/null | grep -A1 "tile_" | grep -v "^--" | grep -o "tile_[a-z]*_kernel\|us *[0-9.]*" | paste - - | head -4
done
