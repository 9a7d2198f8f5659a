#!/bin/bash
# final-state `ncu --set full` of the view kernels (per-view operators: one view forward + backward, second view warm) and of
# the tile-partition kernels of the batch driver
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"composite_bwd_kernel|composite_fwd_kernel|build_sublists|shade_bwd_kernel|shade_fwd_kernel|project_bwd_kernel|project_fwd_kernel" --launch-skip 7 -c 7 -f -o gpurun_out/c59_views python scripts/bench_composite.py --iters 2 > gpurun_out/c59_views.log 2>&1
ls -la gpurun_out/c59_views.ncu-rep
ncu -i gpurun_out/c59_views.ncu-rep --page raw --csv > gpurun_out/c59_views.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c59_views.raw.csv gpurun_out/c59_views.summary.csv 8 | tail -9
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_count_kernel|tile_prefix_kernel|tile_offsets_kernel|tile_scatter_kernel" --launch-skip 64 -c 4 -f -o gpurun_out/c59_tilepart python bench.py --steps 8 --warmup 8 --no-cpu-baseline --no-e2e --no-configs --no-train-step > gpurun_out/c59_tilepart.log 2>&1
ls -la gpurun_out/c59_tilepart.ncu-rep
ncu -i gpurun_out/c59_tilepart.ncu-rep --page raw --csv > gpurun_out/c59_tilepart.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/c59_tilepart.raw.csv gpurun_out/c59_tilepart.summary.csv 4 | tail -5
rm -f gpurun_out/c59_views.raw.csv.tmp
