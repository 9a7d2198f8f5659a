#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in default T96k T192k; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 300 python scripts/bench_prefilter.py 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', [(l['R'], l['plan_fwd_ms'], l['plan_bwd_ms']) for l in d['levels']], d['as_envstack_fwd_bwd_ms'])"
done
unset GSB_LIB_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"specular_apply_kernel" -c 4 -f -o gpurun_out/c30_prof_apply python scripts/bench_prefilter.py > /dev/null 2>&1
ncu -i gpurun_out/c30_prof_apply.ncu-rep --page raw --csv > gpurun_out/c30_prof_apply.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c30_prof_apply.raw.csv gpurun_out/c30_prof_apply.summary.csv 4 | tail -5
