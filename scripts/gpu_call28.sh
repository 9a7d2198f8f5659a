#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/c28_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c28_tests.log
timeout 200 python scripts/bench_train_step.py 140 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train', d['ms_per_step'], d['ms_per_step_min_max'], d['entry_point_ms_one_step'])"
timeout 200 python scripts/profile_train_step.py 140 gpurun_out/c28_train_step_kernels.json | tail -1 | cut -c1-700
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"mlp_bwd_kernel|mlp_fwd_kernel|hashgrid_bwd_lm" -c 6 -f -o gpurun_out/c28_prof_fields python scripts/bench_encoding.py > /dev/null 2>&1
ncu -i gpurun_out/c28_prof_fields.ncu-rep --page raw --csv > gpurun_out/c28_prof_fields.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c28_prof_fields.raw.csv gpurun_out/c28_prof_fields.summary.csv 6 | tail -5
