for ns in 2 4 6; do
  python bench.py --no-cpu-baseline --no-e2e --streams $ns 2>/dev/null | tail -1 > gpurun_out/sweep_$ns.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/sweep_$ns.json')); print('streams', $ns, d['value'], d['ms_per_step'], d['batches']['device_ms'])"
done
python bench.py --no-cpu-baseline --no-e2e --views 16 --streams 4 2>/dev/null | tail -1 > gpurun_out/sweep_b16.json
python -c "
import json,sys; d=json.load(open('gpurun_out/sweep_b16.json')); print('batch16 streams4', d['value'], d['ms_per_step'], d['batches']['device_ms'])"
