#!/bin/bash
for v in PB6 PB8 SF5 SF6; do
export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so
python scripts/bench_composite.py --iters 24 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'project_bwd', d['gsb_project_bwd'], 'shade_fwd', d['gsb_shade_fwd'], d['sum_ms_per_view'])"
done
