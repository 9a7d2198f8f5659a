#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_prefilter_gpu.py -q -x > gpurun_out/c29_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c29_tests.log
for v in default T6k T24k T48k; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 300 python scripts/bench_prefilter.py 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', [(l['R'], l['plan_fwd_ms'], l['plan_bwd_ms']) for l in d['levels']], d['as_envstack_fwd_bwd_ms'])"
done
