#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/c22_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c22_tests.log
timeout 700 python bench.py > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c22_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c22_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'launches', d['gpu_launches'])
print('batches', json.dumps(d['batches'])[:700])
print('e2e', d['e2e']['value'], 'train_step', json.dumps(d.get('train_step'))[:420])
print('configs', json.dumps(d.get('configs'))[:1500])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','avg_ms','avg_ms_alone','frac_alone')})
print('kernels', {k: v['avg_ms'] for k, v in d['kernels'].items()})
print('parity', json.dumps(d.get('parity'))[:800])
print('cpu', d.get('cpu_baseline'))
PY
