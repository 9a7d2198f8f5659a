#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "--- TMA gather4 staging experiment"
timeout 60 ./build/tma_gather_experiment | tee gpurun_out/c11_tma_experiment.json
timeout 120 ncu --set full --clock-control none -k regex:"stage_lanes|stage_tma" -c 6 -f -o gpurun_out/c11_prof_tma ./build/tma_gather_experiment > /dev/null 2>&1
ncu -i gpurun_out/c11_prof_tma.ncu-rep --page raw --csv > gpurun_out/c11_prof_tma.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c11_prof_tma.raw.csv gpurun_out/c11_prof_tma.summary.csv 6
echo "--- tests"
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c11_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c11_tests_gpu.log
echo "--- bench"
timeout 500 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c11_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c11_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'launches', d['gpu_launches'])
print('e2e', json.dumps(d.get('e2e'))[:300])
print('train_step', json.dumps(d.get('train_step'))[:500])
print('kernels', {k: v['avg_ms'] for k, v in d['kernels'].items()})
PY
