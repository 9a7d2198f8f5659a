#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/c6_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c6_tests_gpu.log
timeout 100 python scripts/bench_encoding.py > gpurun_out/c6_encoding.log 2>&1; tail -2 gpurun_out/c6_encoding.log | cut -c1-1200
timeout 120 python scripts/bench_train_step.py 140 > gpurun_out/c6_train_step.log 2>&1; tail -1 gpurun_out/c6_train_step.log | cut -c1-1500
timeout 200 python scripts/profile_train_step.py 140 gpurun_out/c6_train_step_kernels.json > gpurun_out/c6_profile.log 2>&1; tail -1 gpurun_out/c6_profile.log | cut -c1-1500
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"mlp_fwd_kernel|mlp_bwd_kernel|hashgrid_bwd_kernel|hashgrid_fwd_kernel" -c 8 -f \
  -o gpurun_out/c6_prof_fields python scripts/bench_encoding.py > /dev/null 2>&1
ncu -i gpurun_out/c6_prof_fields.ncu-rep --page raw --csv > gpurun_out/c6_prof_fields.raw.csv 2>/dev/null
ls -la gpurun_out/c6_*
