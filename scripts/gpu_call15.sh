#!/bin/bash
# build_sublists at two CTAs per SM, LPT on the longest sub-list; ncu of the non-composite view kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_parity_fullsize_gpu.py tests/test_splat_gpu.py -q -x > gpurun_out/c15_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c15_tests.log
timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
for v in B1 B512; do
  GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
done
timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --steps 80 2>/dev/null | tail -1 > gpurun_out/c15_bench.json
python -c "
import json; d=json.load(open('gpurun_out/c15_bench.json')); k=d['kernels']
print('views/s', d['value'], 'batch ms', sorted(d['batches']['device_ms'])[:3], 'seq', d['sequential_ms_per_view'], 'fwd', k['gsb_composite_fwd']['avg_ms'], 'bwd', k['gsb_composite_bwd']['avg_ms'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"composite_fwd_kernel|composite_bwd_kernel|build_sublists|shade_bwd_kernel|shade_fwd_kernel|project_bwd|isect_tiles|Onesweep|Histogram|replica" -c 30 -f \
  -o gpurun_out/c15_prof_view python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c15_prof_view.ncu-rep --page raw --csv > gpurun_out/c15_prof_view.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c15_prof_view.raw.csv gpurun_out/c15_prof_view.summary.csv 30 | tail -32
