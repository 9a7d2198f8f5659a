#!/bin/bash
# AddressSanitizer + UBSan over the kernels' own source compiled for the host (tests/emu, both modes): every global- and
# shared-memory access of every kernel file in the CPU parity tests is bounds-checked against the numpy / torch
# allocations (or the static shared arrays) it was handed.  A few minutes.
#   bash scripts/memcheck_host.sh [pytest args]      (log: profiles/r02_host_asan.log)
set -o pipefail
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
GSB_EMU_SANITIZE=1 LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_project_cpu.py tests/test_flexicubes_cpu.py tests/test_shade_cpu.py tests/test_fields_cpu.py \
  tests/test_composite_cpu.py tests/test_loss_cpu.py tests/test_prefilter_cpu.py tests/test_view_driver_cpu.py \
  tests/test_edge_cases_cpu.py \
  -q -p no:cacheprovider "$@" 2>&1 | tee profiles/r02_host_asan.log | tail -5
