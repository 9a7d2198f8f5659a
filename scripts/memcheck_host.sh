#!/bin/bash
# AddressSanitizer + UBSan over the streaming kernels' own source compiled for the host (tests/emu, sequential mode):
# every global-memory access of project / binsort / shade / texture / hashgrid / mgadapter / flexicubes kernels in the
# CPU parity tests is bounds-checked against the numpy / torch allocations it was handed.  ~3 minutes.
#   bash scripts/memcheck_host.sh [pytest args]      (log: profiles/r01_host_asan.log)
set -o pipefail
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
GSB_EMU_SANITIZE=1 LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_project_cpu.py tests/test_flexicubes_cpu.py tests/test_shade_cpu.py tests/test_fields_cpu.py \
  -q -p no:cacheprovider -k "not simt" "$@" 2>&1 | tee profiles/r01_host_asan.log | tail -5
