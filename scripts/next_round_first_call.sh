#!/bin/bash
# What to run first on a B200 next round (one gpurun call, ~6 min of box time): the parity gate, the bench line, and the
# three ncu captures DESIGN.md section 10 asks for before touching a kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 600 -- 'bash scripts/next_round_first_call.sh'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests -q -m gpu > gpurun_out/r2_tests_gpu.log 2>&1; echo "pytest rc=$?"
timeout 150 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 60 python scripts/bench_train_step.py 140 > gpurun_out/r2_train_step.log 2>&1; echo "train step rc=$?"
# launch list of the bench command (shares, cold-cache) -- profiles/r02_launches.csv
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 8 --warmup 8 --no-cpu-baseline --no-e2e --no-train-step > /dev/null 2>&1; echo "launch list rc=$?"
# full-set captures: prefilter gather (lever 2), hash-grid backward (lever 3), composite backward (lever 1)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:specular_gather -c 12 -f \
  -o gpurun_out/r2_prof_prefilter python scripts/bench_prefilter.py > /dev/null 2>&1
ncu -i gpurun_out/r2_prof_prefilter.ncu-rep --page raw --csv > gpurun_out/r2_prof_prefilter.raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"hashgrid_bwd|composite_bwd" -c 6 -f \
  -o gpurun_out/r2_prof_bwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-train-step --full-step \
  > /dev/null 2>&1
ncu -i gpurun_out/r2_prof_bwd.ncu-rep --page raw --csv > gpurun_out/r2_prof_bwd.raw.csv 2>/dev/null
ls -la gpurun_out/r2_* | head -20
