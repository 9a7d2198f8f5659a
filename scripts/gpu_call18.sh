#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_encoding_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/c18_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c18_tests.log
timeout 200 python scripts/bench_encoding.py | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())['fields']; print('default', {k:(v['encode_fwd_ms'], v['encode_fwd_bwd_ms']) for k,v in d.items()})"
for v in A0 A3 A12 A32; do
GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so timeout 200 python scripts/bench_encoding.py | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read())['fields']; print('$v', {k:(v['encode_fwd_ms'], v['encode_fwd_bwd_ms']) for k,v in d.items()})"
done
timeout 200 python scripts/bench_train_step.py 140 | tail -1 | cut -c1-600
GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_A0.so timeout 200 python scripts/bench_train_step.py 140 | tail -1 | cut -c1-300
