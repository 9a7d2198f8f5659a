#!/bin/bash
for v in default PB4 PB5; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
python scripts/bench_composite.py --iters 24 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['gsb_project_bwd'], d['sum_ms_per_view'])"
done
