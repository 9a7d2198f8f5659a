#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_model_gpu.py tests/test_flexicubes_gpu.py -q -x > gpurun_out/c34_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c34_tests.log
timeout 200 python scripts/bench_train_step.py 140 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train', d['ms_per_step'], d['ms_per_step_min_max'])"
timeout 200 python scripts/profile_train_step.py 140 gpurun_out/c34_train_step_kernels.json | tail -1 | cut -c1-600
python - <<'PY'
import json
d=json.load(open('gpurun_out/c34_train_step_kernels.json'))
for o in d['torch_ops_by_self_device_time'][:14]: print(o['op'][:50], o['calls'], o['ms'], o['shapes'][:60])
PY
