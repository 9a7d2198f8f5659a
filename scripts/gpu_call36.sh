#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_raster_gpu.py tests/test_parity_fullsize_gpu.py tests/test_splat_gpu.py -q -x > gpurun_out/c36_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c36_tests.log
for v in default OLD default OLD; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 100 python scripts/bench_composite.py --iters 24 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['gsb_composite_fwd'], d['gsb_composite_bwd'], d['sum_ms_per_view'])"
done
for v in default OLD; do
if [ $v = default ]; then unset GSB_LIB_PATH; else export GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so; fi
timeout 300 python bench.py --no-cpu-baseline --no-configs --no-e2e --steps 80 2>/dev/null | tail -1 > gpurun_out/c36_bench_$v.json
python -c "
import json; d=json.load(open('gpurun_out/c36_bench_$v.json')); k=d['kernels']
print('$v', 'views/s', d['value'], 'batch ms', sorted(d['batches']['device_ms'])[:3], 'bwd', k['gsb_composite_bwd']['avg_ms'])"
done
unset GSB_LIB_PATH
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"composite_bwd_kernel" -c 1 -f -o gpurun_out/c36_prof_bwd python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c36_prof_bwd.ncu-rep --page raw --csv > gpurun_out/c36_prof_bwd.raw.csv 2>/dev/null
python scripts/ncu_pick.py gpurun_out/c36_prof_bwd.raw.csv gpu__time_duration.sum smsp__inst_executed.sum l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum smsp__issue_active.avg.pct_of_peak_sustained_active launch__registers_per_thread
