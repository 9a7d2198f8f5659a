#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_encoding_gpu.py tests/test_field_gpu.py -q -x > gpurun_out/c26_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c26_tests.log
for v in default P8W7 P4W14 P4W12 P4W8; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 200 python scripts/bench_mlp.py | tail -1
done
