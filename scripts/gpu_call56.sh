#!/bin/bash
# 4-GPU check of both bench arms as the driver launches them (N = 4 was never run this round)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/c56_bench_n4.json 2> gpurun_out/c56_bench_n4.err; echo "bench n4 rc=$?"; tail -3 gpurun_out/c56_bench_n4.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/c56_ref_n4.json 2> gpurun_out/c56_ref_n4.err; echo "ref n4 rc=$?"; tail -2 gpurun_out/c56_ref_n4.err; tail -c 600 gpurun_out/c56_ref_n4.json
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c56_bench_n4.json').read().strip().splitlines()[-1])
    print('N=4 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'parallelism', d['config'].get('parallelism'), 'clocks', d.get('clocks'))
    print('configs', json.dumps(d.get('configs')))
except Exception as e: print('parse failed', e)
PY
