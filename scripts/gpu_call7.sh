#!/bin/bash
# 2-GPU check of the bench's N > 1 path (flat all-reduce, configs) + first-call latency of the new bounds kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 60 python - <<'PY' 2>&1 | tail -3
import time, torch, sys
sys.path.insert(0, '.')
from geosplatting_b200 import splitsum
torch.cuda.init(); dev='cuda:0'
for R, r in ((512, 0.08), (256, 0.185), (128, 0.29), (64, 0.395), (32, 0.5), (16, 1.0)):
    ct = splitsum.ndf_cutoff_costheta(r, 0.99)
    torch.cuda.synchronize(); t0=time.perf_counter()
    b = splitsum.render_utils.specular_bounds(R, ct, 0)
    torch.cuda.synchronize(); print('bounds', R, r, round((time.perf_counter()-t0)*1e3, 3), 'ms', 'mean box area', float(((b[...,1::4]-b[...,0::4]+1).clamp_min(0)*(b[...,3::4]-b[...,2::4]+1).clamp_min(0)).sum(-1).mean()))
PY
timeout 300 python -m pytest tests/test_prefilter_gpu.py -q > gpurun_out/c7_prefilter_tests.log 2>&1; tail -2 gpurun_out/c7_prefilter_tests.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 48 --warmup 8 > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/c7_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c7_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'parallelism', d['config']['parallelism'])
    print('configs', json.dumps(d.get('configs')))
except Exception as e: print('parse failed', e)
PY
