#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c12_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c12_tests_gpu.log
timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
timeout 600 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c12_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c12_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'launches', d['gpu_launches'])
print('batches', json.dumps(d['batches'])[:900])
print('e2e', d['e2e']['value'], 'train_step', json.dumps(d.get('train_step'))[:420])
print('configs', json.dumps(d.get('configs'))[:1500])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','avg_ms','avg_ms_alone','frac_alone')})
PY
timeout 120 ncu --set full --clock-control none -k regex:"stage_tma" -c 2 -f -o gpurun_out/c12_prof_tma ./build/tma_gather_experiment > /dev/null 2>&1
ncu -i gpurun_out/c12_prof_tma.ncu-rep --page raw --csv > gpurun_out/c12_prof_tma.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c12_prof_tma.raw.csv gpurun_out/c12_prof_tma.summary.csv 2
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"hashgrid_bwd_lm|mlp_bwd_kernel|mlp_fwd_kernel|hashgrid_fwd_lm" -c 8 -f \
  -o gpurun_out/c12_prof_fields python scripts/bench_encoding.py > /dev/null 2>&1
ncu -i gpurun_out/c12_prof_fields.ncu-rep --page raw --csv > gpurun_out/c12_prof_fields.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c12_prof_fields.raw.csv gpurun_out/c12_prof_fields.summary.csv 8
timeout 100 python scripts/profile_train_step.py 140 gpurun_out/c12_train_step_kernels.json | tail -1 | cut -c1-900
