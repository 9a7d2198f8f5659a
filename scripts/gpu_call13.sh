#!/bin/bash
# segmented persistent compositing backward: parity gate, per-entry-point timings of the default and the tuning builds, bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_parity_fullsize_gpu.py tests/test_splat_gpu.py -q -x > gpurun_out/c13_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c13_tests.log
timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
for v in SEG64 SEG128 SEG512 C3 C4 C8; do
  GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
done
timeout 300 python bench.py --no-cpu-baseline --no-configs > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c13_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c13_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'launches', d['gpu_launches'])
print('batches', json.dumps(d['batches'])[:600])
print('e2e', d['e2e']['value'], 'train_step', json.dumps(d.get('train_step'))[:300])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','avg_ms','avg_ms_alone','frac_alone')})
print('parity', json.dumps(d.get('parity'))[:600])
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"composite_bwd_kernel|composite_fwd_kernel|build_sublists" -c 3 -f \
  -o gpurun_out/c13_prof_comp python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c13_prof_comp.ncu-rep --page raw --csv > gpurun_out/c13_prof_comp.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c13_prof_comp.raw.csv gpurun_out/c13_prof_comp.summary.csv 3
