#!/bin/bash
# end-of-round verification: the whole GPU suite, smoke(), the default bench line, the reference arm, the launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c37_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c37_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/c37_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'launches', d['gpu_launches'], 'e2e', d['e2e']['value'])
print('train_step', json.dumps(d.get('train_step'))[:400])
print('configs', {k: (v.get('views_per_s') or v.get('D14_ms_per_view')) for k, v in d.get('configs', {}).items()})
print('roofline', d['roofline'])
print('parity', json.dumps(d.get('parity'))[:500])
print('clocks', d.get('clocks'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c37_launches.csv python bench.py --steps 8 --warmup 8 --no-cpu-baseline --no-e2e --no-configs --no-train-step > /dev/null 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/c37_launches.csv
