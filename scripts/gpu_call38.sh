#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_encoding_gpu.py tests/test_field_gpu.py -q -x > gpurun_out/c38_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c38_tests.log
for v in default OLD K1 K2 K6 K8; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 200 python scripts/bench_encoding.py | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())['fields']; print('$v', {k:(v['encode_fwd_ms'], v['encode_fwd_bwd_ms']) for k,v in d.items()})"
done
for v in default OLD; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 200 python scripts/bench_train_step.py 140 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['ms_per_step_min_max'], d['entry_point_ms_one_step'].get('gsb_hashgrid_bwd'))"
done
