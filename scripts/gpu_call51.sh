#!/bin/bash
for v in default CF5 CF6 SB4 SB2; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
python scripts/bench_composite.py --iters 24 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'composite_fwd', d['gsb_composite_fwd'], 'shade_bwd', d['gsb_shade_bwd'], 'project_bwd', d['gsb_project_bwd'], 'shade_fwd', d['gsb_shade_fwd'], d['sum_ms_per_view'])"
done
