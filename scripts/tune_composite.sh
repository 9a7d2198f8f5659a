for v in base FG4 FG16 BG4 BG16 WB1 W4 W1; do
  if [ $v = base ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
  python bench.py --no-cpu-baseline --no-e2e --steps 40 2>/dev/null | tail -1 > gpurun_out/tune_$v.json
  python -c "
import json; d=json.load(open('gpurun_out/tune_$v.json')); k=d['kernels']
print('$v', 'views/s', d['value'], 'steady batch ms', sorted(d['batches']['device_ms'])[1], 'seq ms/view', d['sequential_ms_per_view'], 'fwd', k['gsb_composite_fwd']['avg_ms'], 'bwd', k['gsb_composite_bwd']['avg_ms'])"
done
