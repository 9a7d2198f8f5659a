#!/bin/bash
# round 2, call 3: two-units-per-warp composite + batch driver: parity, per-kernel timings, bench, ncu
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c3_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c3_tests_gpu.log
timeout 120 python scripts/bench_composite.py > gpurun_out/c3_comp_base.json 2> gpurun_out/c3_comp_base.err; cat gpurun_out/c3_comp_base.json
timeout 300 python bench.py --steps 80 --no-train-step > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c3_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'e2e', d['e2e']['value'] if d.get('e2e') else None)
print('batches', d['batches'])
print('roofline', d['roofline']['avg_ms'], d['roofline']['avg_ms_alone'])
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"composite_fwd_kernel|composite_bwd_kernel|build_sublists" -c 6 -f \
  -o gpurun_out/c3_prof_composite python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c3_prof_composite.ncu-rep --page raw --csv > gpurun_out/c3_prof_composite.raw.csv 2>/dev/null
ls -la gpurun_out/c3_* | head -30
