#!/bin/bash
# float2 slab + occupancy variants of the segmented backward
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_parity_fullsize_gpu.py tests/test_splat_gpu.py -q -x -s > gpurun_out/c14_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c14_tests.log; grep "^parity" gpurun_out/c14_tests.log | cut -c1-900
timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
for v in M4 M5 M6 M4S128 M4BG8; do
  GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
done
for v in base M4 M5; do
  if [ $v = base ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --steps 80 2>/dev/null | tail -1 > gpurun_out/c14_bench_$v.json
  python -c "
import json; d=json.load(open('gpurun_out/c14_bench_$v.json')); k=d['kernels']
print('$v', 'views/s', d['value'], 'batch ms', sorted(d['batches']['device_ms'])[:3], 'seq', d['sequential_ms_per_view'], 'fwd', k['gsb_composite_fwd']['avg_ms'], 'bwd', k['gsb_composite_bwd']['avg_ms'])"
done
unset GSB_LIB_PATH
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"composite_bwd_kernel" -c 1 -f \
  -o gpurun_out/c14_prof_comp python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c14_prof_comp.ncu-rep --page raw --csv > gpurun_out/c14_prof_comp.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c14_prof_comp.raw.csv gpurun_out/c14_prof_comp.summary.csv 3
