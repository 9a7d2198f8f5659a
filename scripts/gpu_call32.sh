#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in default D2 D3 D6 D4M10 D4M12 D6M8; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 300 python scripts/bench_prefilter.py 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', [(l['R'], l['plan_fwd_ms'], l['plan_stream_gbs']) for l in d['levels'][:4]], d['as_envstack_fwd_bwd_ms'])"
done
