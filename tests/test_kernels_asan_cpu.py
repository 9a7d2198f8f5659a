"""AddressSanitizer + UBSan over kernel source compiled for the host (tests/emu, GSB_EMU_SANITIZE=1): the projection,
tile-emission and binning kernels' global-memory accesses are bounds-checked against the buffers the parity tests hand
them, in a subprocess with libasan preloaded.  scripts/memcheck_host.sh runs the same over EVERY emulation test, SIMT
mode included (8 minutes; log in profiles/r01_host_asan.log); a negative control shows the checker sees a kernel's
stray write."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _asan_env():
    try:
        lib = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
    except Exception:
        lib = ""
    if not lib or not os.path.exists(lib):
        pytest.skip("libasan not available")
    env = dict(os.environ, GSB_EMU_SANITIZE="1", LD_PRELOAD=lib, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    # the sanitizer runtime needs ~20 TB of address space for its shadow: skip where the environment does not allow it
    probe = subprocess.run([sys.executable, "-c", "import numpy, ctypes; print('asan-ok')"], env=env, capture_output=True,
                           text=True, timeout=300)
    if probe.returncode != 0 or "asan-ok" not in probe.stdout:
        pytest.skip("python does not start under a preloaded libasan here: " + probe.stderr[-200:])
    return env


def test_projection_and_binning_kernels_are_clean_under_asan():
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_project_cpu.py", "-q", "-x", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=_asan_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr + r.stdout and "runtime error" not in r.stderr, \
        (r.stdout[-2000:], r.stderr[-2000:])


def test_the_checker_sees_a_stray_kernel_write():
    code = (
        "import ctypes as C, numpy as np, sys\n"
        "sys.path.insert(0, '.')\n"
        "from tests.emu import build as emu\n"
        "lib = emu.build('flexicubes')\n"
        "F = 64\n"
        "sdf = (np.random.rand(27) - 0.5).astype(np.float32)\n"
        "cubes = np.random.randint(0, 27, (F, 8)).astype(np.int32)\n"
        "cases, flag = np.zeros(F, np.int32), np.zeros(F - 8, np.int32)      # surf_flag 8 entries short\n"
        "p = lambda a: C.c_void_p(a.ctypes.data)\n"
        "lib.gsb_fc_classify(C.c_int32(F), p(sdf), p(cubes), p(cases), p(flag), None)\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=_asan_env(), capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "heap-buffer-overflow" in r.stderr and "fc_classify_kernel" in r.stderr
