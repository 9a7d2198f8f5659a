#!/bin/bash
# round 2, call 2: new composite kernels -- parity (incl. full-size oracle tests), per-kernel timings, tuning sweep, ncu
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -q -m gpu -x > gpurun_out/c2_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c2_tests_gpu.log
timeout 120 python scripts/bench_composite.py > gpurun_out/c2_comp_base.json 2> gpurun_out/c2_comp_base.err; cat gpurun_out/c2_comp_base.json
for v in W2 W8 FG4 FG16 BG4 BG16 WB2; do
  GSB_LIB_PATH=$PWD/geosplatting_b200/lib/tune_$v.so timeout 100 python scripts/bench_composite.py --iters 12 2>/dev/null | tail -1 | tee gpurun_out/c2_comp_$v.json
done
timeout 200 python bench.py --steps 40 --no-train-step > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"composite_fwd_kernel|composite_bwd_kernel|build_sublists" -c 6 -f \
  -o gpurun_out/c2_prof_composite python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ncu -i gpurun_out/c2_prof_composite.ncu-rep --page raw --csv > gpurun_out/c2_prof_composite.raw.csv 2>/dev/null
ls -la gpurun_out/c2_* | head -30
