#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "--- TMA gather4 staging experiment"
timeout 60 ./build/tma_gather_experiment | tee gpurun_out/c10_tma_experiment.json
timeout 120 ncu --set full --clock-control none -k regex:"stage_lanes|stage_tma" -c 4 -f -o gpurun_out/c10_prof_tma ./build/tma_gather_experiment > /dev/null 2>&1
ncu -i gpurun_out/c10_prof_tma.ncu-rep --page raw --csv > gpurun_out/c10_prof_tma.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c10_prof_tma.raw.csv gpurun_out/c10_prof_tma.summary.csv 4
echo "--- bench"
timeout 500 python bench.py > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/c10_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'seq', d['sequential_ms_per_view'], 'e2e', d['e2e']['value'] if d.get('e2e') else None, 'launches', d['gpu_launches'])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','avg_ms','avg_ms_alone','frac_alone')})
print('train_step', json.dumps(d.get('train_step'))[:400])
print('e2e_training_view', json.dumps(d.get('e2e_training_view'))[:200])
print('kernels', {k: v['avg_ms'] for k, v in d['kernels'].items()})
PY
echo "--- ncu launch list + captures"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/c10_launches.csv \
  python bench.py --steps 8 --warmup 8 --no-cpu-baseline --no-e2e --no-train-step --no-configs > /dev/null 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"composite_fwd_kernel|composite_bwd_kernel|build_sublists|shade_bwd_kernel" -c 8 -f \
  -o gpurun_out/c10_prof_views python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c10_prof_views.ncu-rep --page raw --csv > gpurun_out/c10_prof_views.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c10_prof_views.raw.csv gpurun_out/c10_prof_views.summary.csv 8
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"hashgrid_bwd_lm|mlp_bwd_kernel|specular_apply" -c 6 -f \
  -o gpurun_out/c10_prof_step python scripts/bench_train_step.py 140 > /dev/null 2>&1
ncu -i gpurun_out/c10_prof_step.ncu-rep --page raw --csv > gpurun_out/c10_prof_step.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c10_prof_step.raw.csv gpurun_out/c10_prof_step.summary.csv 6
