#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_encoding_gpu.py tests/test_field_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/c25_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c25_tests.log
for v in default OLD W4 W7; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 200 python scripts/bench_encoding.py | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())['fields']; print('$v', {k:(v['encode_fwd_bwd_ms'], v.get('field_fwd_bwd_ms')) for k,v in d.items()})"
done
for v in default OLD; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 200 python scripts/bench_train_step.py 140 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['ms_per_step_min_max'], d['entry_point_ms_one_step'].get('gsb_mlp_fwd'), d['entry_point_ms_one_step'].get('gsb_mlp_bwd'))"
done
unset GSB_LIB_PATH
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"mlp_bwd_kernel|mlp_fwd_kernel" -c 4 -f -o gpurun_out/c25_prof_mlp python scripts/bench_encoding.py > /dev/null 2>&1
ncu -i gpurun_out/c25_prof_mlp.ncu-rep --page raw --csv > gpurun_out/c25_prof_mlp.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c25_prof_mlp.raw.csv gpurun_out/c25_prof_mlp.summary.csv 4 | tail -5
