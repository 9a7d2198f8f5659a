#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_encoding_gpu.py -q -x > gpurun_out/c19_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c19_tests.log
timeout 200 python scripts/bench_train_step.py 140 > /dev/null 2>&1
timeout 200 python scripts/bench_encoding.py | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())['fields']; print('default', {k:(v['encode_fwd_ms'], v['encode_fwd_bwd_ms']) for k,v in d.items()})"
for v in A0 A2 A8 A16; do
GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so timeout 200 python scripts/bench_encoding.py | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())['fields']; print('$v', {k:(v['encode_fwd_ms'], v['encode_fwd_bwd_ms']) for k,v in d.items()})"
done
for v in default A0 A8 default; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 200 python scripts/bench_train_step.py 140 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['entry_point_ms_one_step'].get('gsb_hashgrid_bwd'), d['entry_point_ms_one_step'].get('gsb_mlp_bwd'))"
done
