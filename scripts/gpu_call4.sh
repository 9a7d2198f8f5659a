#!/bin/bash
# round 2, call 4: pipelined gathers + float4 slab; stream-count sweep; parity block + other configs in bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c4_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c4_tests_gpu.log
timeout 120 python scripts/bench_composite.py > gpurun_out/c4_comp_base.json 2> gpurun_out/c4_comp_base.err; cat gpurun_out/c4_comp_base.json
for ns in 2 4 6 8; do
  timeout 200 python bench.py --steps 80 --streams $ns --no-train-step --no-cpu-baseline --no-e2e --no-configs > gpurun_out/c4_bench_s$ns.json 2> gpurun_out/c4_bench_s$ns.err
  python - <<PY
import json
d=json.load(open('gpurun_out/c4_bench_s$ns.json'))
print('streams', $ns, 'value', d['value'], 'ms/step', d['ms_per_step'], 'host', sorted(d['batches']['host_enqueue_ms'])[len(d['batches']['host_enqueue_ms'])//2], 'dev', sorted(d['batches']['device_ms'])[len(d['batches']['device_ms'])//2], 'bwd live', d['roofline']['avg_ms'])
PY
done
timeout 400 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c4_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'e2e', d['e2e']['value'] if d.get('e2e') else None)
print('parity', json.dumps(d.get('parity')))
print('configs', json.dumps(d.get('configs')))
print('train_step', json.dumps(d.get('train_step'))[:300])
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"composite_fwd_kernel|composite_bwd_kernel" -c 4 -f \
  -o gpurun_out/c4_prof_composite python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c4_prof_composite.ncu-rep --page raw --csv > gpurun_out/c4_prof_composite.raw.csv 2>/dev/null
ls -la gpurun_out/c4_* | head -30
