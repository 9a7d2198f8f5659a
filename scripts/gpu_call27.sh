#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in default B10 B11 B13 U8 U2 F22; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 200 python scripts/bench_mlp.py | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', {k:(v['fwd_ms'], v['bwd_ms']) for k,v in d.items() if k!='lib'})"
done
