#!/bin/bash
# compute-sanitizer over the kernels that changed this round (persistent segmented backward, checkpoints and job queue,
# tiled MLPs in opt-in shared memory, hash-grid warp aggregation, plan ring buffer, packing shade forward)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5"
( timeout 900 $S python -m pytest -q -x "tests/test_raster_gpu.py::test_config1_10k_256" "tests/test_raster_gpu.py::test_ragged_resolution_and_big_splats" "tests/test_raster_gpu.py::test_wide_channel_backward" "tests/test_splat_gpu.py::test_multi_stream_batch_equals_sequential_views" "tests/test_splat_gpu.py::test_batch_edge_cases_empty_scene_and_mixed_resolutions" "tests/test_encoding_gpu.py::test_clustered_points_table_gradient_matches_oracle" "tests/test_encoding_gpu.py::test_edge_cases" "tests/test_encoding_gpu.py::test_fields_match_reference_code" "tests/test_prefilter_gpu.py::test_as_splitsum_and_envstack_agree_and_are_finite" ; echo "memcheck rc=$?" ) > gpurun_out/c35_memcheck.log 2>&1
tail -4 gpurun_out/c35_memcheck.log; grep -c "Invalid\|Error:" gpurun_out/c35_memcheck.log
S2="compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5"
( timeout 600 $S2 python -m pytest -q -x "tests/test_raster_gpu.py::test_config1_10k_256" "tests/test_encoding_gpu.py::test_fields_match_reference_code" ; echo "racecheck rc=$?" ) > gpurun_out/c35_racecheck.log 2>&1
tail -4 gpurun_out/c35_racecheck.log
