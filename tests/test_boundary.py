"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

from geosplatting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.gsb_version.restype = ctypes.c_int
    assert lib.gsb_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.gsb_bin_workspace_bytes(ctypes.c_int32(-1), ctypes.c_int64(0), None)
    assert rc == -1
    assert b"invalid argument" in lib.gsb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "geosplatting_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "liboracle" in src \
                        or re.search(r'#include\s+".*oracle', src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_cpu_tensors_are_rejected():
    import torch
    from geosplatting_b200 import rasterization
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rasterization(z, torch.zeros(4, 4), z, torch.zeros(4), z, torch.eye(4)[None], torch.eye(3)[None], 32, 32)


def test_header_is_plain_c_and_every_error_path_answers_without_a_gpu():
    """include/geosplat_b200.h compiles as C99 (no torch / C++ types in the ABI), and argument validation of the newer
    entry points (two-stage binning, per-view driver, hash grid, loss) answers with GSB_EINVAL before touching a device."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc:
        subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c",
                               _lib.HEADER_PATH])
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    assert lib.gsb_bin2_workspace_bytes(ctypes.c_int32(-1), ctypes.c_int64(0), ctypes.byref(n)) == -1
    assert lib.gsb_bin2_workspace_bytes(ctypes.c_int32(1000), ctypes.c_int64(5000), ctypes.byref(n)) == 0 and n.value > 0
    assert lib.gsb_view_bytes(None, ctypes.c_int64(0), None) == -1
    cfg = _lib.GsbViewConfig(1000, 64, 48, 256, 64, 3, 16, 0.1, 1.0, 0.08, 0.5, 0, 1)
    sizes = (ctypes.c_size_t * 5)()
    assert lib.gsb_view_bytes(ctypes.addressof(cfg), 4000, ctypes.addressof(sizes)) == 0 and all(s > 0 for s in sizes)
    bad = _lib.GsbViewConfig(1000, 64, 48, 256, 64, 3, 16, 0.1, 1.0, 0.08, 0.5, 7, 1)      # mode out of range
    assert lib.gsb_view_bytes(ctypes.addressof(bad), 0, ctypes.addressof(sizes)) == -1
    sc = (ctypes.c_float * 16)(*range(16, 32))
    assert lib.gsb_hashgrid_fwd(ctypes.c_int64(10), None, None, ctypes.c_int32(16), ctypes.c_int32(4), ctypes.c_int32(18),
                                sc, None, None) == -1
    assert b"features_per_level must be 2" in lib.gsb_last_error()
    assert lib.gsb_loss_fwd(ctypes.c_int32(8), ctypes.c_int32(8), None, None, None, ctypes.c_float(0.2),
                            ctypes.c_float(5.0), None, None, None) == -1
    assert lib.gsb_shade_workspace_bytes(ctypes.c_int32(512), ctypes.c_int32(6), ctypes.c_int32(16), ctypes.byref(n)) == 0
    assert n.value == 16 * (6 * 32 * 32 + 6 * 16 * 16 + 6 * 16 * 16) * 32


def test_flexicubes_entry_points_validate_arguments_without_a_gpu():
    """gsb_fc_*: size queries answer on the host, bad arguments come back as GSB_EINVAL with a message."""
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    i32, i64 = ctypes.c_int32, ctypes.c_int64
    assert lib.gsb_fc_workspace_bytes(i32(128 ** 3), i32(0), ctypes.byref(n)) == 0 and n.value > 0
    small = n.value
    assert lib.gsb_fc_workspace_bytes(i32(128 ** 3), i32(100000), ctypes.byref(n)) == 0 and n.value > small
    # 12 N (number of cube edges) must fit the int32 the sort and the scans count in; more surface cubes than cubes
    assert lib.gsb_fc_workspace_bytes(i32(2 ** 31 - 1), i32(2 ** 28), ctypes.byref(n)) == -1
    assert lib.gsb_fc_workspace_bytes(i32(10), i32(11), ctypes.byref(n)) == -1
    cnt = i32(7)
    assert lib.gsb_fc_surface(i32(0), None, None, None, None, None, None, ctypes.c_size_t(0), ctypes.byref(cnt), None) == 0
    assert cnt.value == 0                                                     # empty grid: no device work, N = 0
    assert lib.gsb_fc_surface(i32(8), None, None, None, None, None, None, ctypes.c_size_t(0), ctypes.byref(cnt), None) == -1
    assert b"gsb_fc_surface" in lib.gsb_last_error()
    counts = (i32 * 4)()
    assert lib.gsb_fc_topology(i32(8), i32(0), i64(27), i32(2), i32(2), i32(2), *([None] * 17), ctypes.c_size_t(0),
                               counts, None) == -1
    assert lib.gsb_fc_quad_gather(i32(0), None, None, None, None) == 0
    assert lib.gsb_fc_quad_gather(i32(4), None, None, None, None) == -1
    assert lib.gsb_fc_dual_fwd(i32(4), *([None] * 18), None) == -1
    assert lib.gsb_fc_entropy_fwd(i64(-1), None, None, None, None, None) == -1
