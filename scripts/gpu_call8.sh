#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python scripts/bench_prefilter.py > gpurun_out/c8_prefilter.log 2>&1; tail -8 gpurun_out/c8_prefilter.log | cut -c1-400
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c8_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c8_tests_gpu.log
timeout 120 python scripts/bench_train_step.py 140 > gpurun_out/c8_train_step.log 2>&1; tail -1 gpurun_out/c8_train_step.log | cut -c1-1500
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"specular_apply_kernel" -c 4 -f \
  -o gpurun_out/c8_prof_apply python scripts/bench_prefilter.py > /dev/null 2>&1
ncu -i gpurun_out/c8_prof_apply.ncu-rep --page raw --csv > gpurun_out/c8_prof_apply.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c8_prof_apply.raw.csv gpurun_out/c8_prof_apply.summary.csv 4
