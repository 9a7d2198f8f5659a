#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python scripts/bench_prefilter.py > gpurun_out/c9_prefilter.log 2>&1; tail -8 gpurun_out/c9_prefilter.log | cut -c1-420
echo "--- hashgrid level-major"; timeout 100 python scripts/bench_encoding.py 2>/dev/null | tail -1 | cut -c1-900
echo "--- hashgrid point-major"; GSB_HASHGRID_POINT_MAJOR=1 timeout 100 python scripts/bench_encoding.py 2>/dev/null | tail -1 | cut -c1-900
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c9_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c9_tests_gpu.log
timeout 120 python scripts/bench_train_step.py 140 > gpurun_out/c9_train_step.log 2>&1; tail -1 gpurun_out/c9_train_step.log | cut -c1-1500
GSB_PREFILTER_PLAN_GB=0 timeout 120 python scripts/bench_train_step.py 140 2>/dev/null | tail -1 | cut -c1-300
