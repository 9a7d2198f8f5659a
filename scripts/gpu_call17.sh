#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
./build/cub_policy_experiment > gpurun_out/c17_cub_policy.json; cat gpurun_out/c17_cub_policy.json
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/c17_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c17_tests.log
timeout 100 python scripts/bench_composite.py --iters 16 | tail -1
