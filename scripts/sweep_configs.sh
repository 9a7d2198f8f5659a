# BASELINE.json configs 2-5 on one GPU: views/s (fwd+bwd, batches of 8 over 3 streams) and the sequential per-view time
for cfg in "83 800" "118 800" "167 800" "264 1600"; do
  set -- $cfg
  python bench.py --mesh-n $1 --res $2 --steps 40 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/cfg_$1_$2.json
  python -c "
import json; d=json.load(open('gpurun_out/cfg_$1_$2.json')); c=d['config']
print('N', c['gaussians'], 'res', c['resolution'][0], 'M', c['intersections'], 'views/s', d['value'], 'ms/view', d['ms_per_step'], 'sequential ms/view', d['sequential_ms_per_view'], 'composite fwd/bwd ms', d['kernels']['gsb_composite_fwd']['avg_ms'], d['kernels']['gsb_composite_bwd']['avg_ms'])"
done
