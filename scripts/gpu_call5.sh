#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in base FG16 FG4 BG16 BG4 W2 W8 WB2 WB2M12 WB4M5 WB4M6 W4M7 FG16BG16; do
  if [ $v = base ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
  timeout 100 python scripts/bench_composite.py --iters 16 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'fwd', d['gsb_composite_fwd'], 'bwd', d['gsb_composite_bwd'], 'sum', d['sum_ms_per_view'])"
done
unset GSB_LIB_PATH
timeout 120 python scripts/bench_train_step.py 140 > gpurun_out/c5_train_step.log 2>&1; tail -1 gpurun_out/c5_train_step.log | cut -c1-1500
timeout 200 python scripts/profile_train_step.py 140 gpurun_out/c5_train_step_kernels.json > gpurun_out/c5_profile.log 2>&1; tail -2 gpurun_out/c5_profile.log | cut -c1-1500
